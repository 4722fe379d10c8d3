"""World-size-2 tests of the multi-tower path (the reference's in-graph towers, TFModel.py:253-377,
VarNetUtility.py:830-857,900-901): contiguous test-function ranges per rank, BC/IC rows replicated and
down-weighted, gradients/losses summed by one all-reduce, weights updated identically on every rank.
CPU: gloo + the oracle-backed test engine.  GPU: nccl + the real engine under torchrun (needs 2 GPUs)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cpu_world(world, steps, batchNum, port):
    import torch.multiprocessing as mp
    from tests import mp_worker
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "w%d.npz" % world)
        os.environ["MASTER_PORT"] = str(port)
        if world == 1:
            mp_worker.run("cpu", out, steps, 1, 0, batchNum)
        else:
            mp.spawn(mp_worker._spawned, args=(world, out, steps, batchNum), nprocs=world, join=True)
        z = np.load(out)
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("batchNum", [None, 2])
def test_two_ranks_equal_one_rank_on_cpu(batchNum):
    one = _cpu_world(1, 3, batchNum, 29541)
    two = _cpu_world(2, 3, batchNum, 29542 if batchNum is None else 29543)
    assert np.allclose(one["split"], two["split"], rtol=1e-12)
    assert np.allclose(one["lossVec"], two["lossVec"], rtol=1e-12, atol=0)
    if batchNum is None:
        # same global batch, split over two towers: identical trajectory up to summation order
        assert np.allclose(one["losses"], two["losses"], rtol=1e-10)
        assert np.allclose(one["theta"], two["theta"], rtol=1e-5, atol=1e-7)
    else:
        # with mini-batches the reference slices batchLen = ceil(nt/batchNum/puNum) per tower, so the
        # mini-batch composition differs between 1 and 2 towers; both must still descend
        assert two["losses"][-1] < two["losses"][0] and one["losses"][-1] < one["losses"][0]


@pytest.mark.gpu
def test_two_gpus_equal_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for world in (1, 2):
            out = os.path.join(d, "g%d.npz" % world)
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                   "--master-addr", "127.0.0.1", "--master-port", str(29551 + world),
                   os.path.join(ROOT, "tests", "mp_worker.py"), "gpu", out, "4"]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
            assert r.returncode == 0, r.stderr[-2000:]
            z = np.load(out)
            outs.append({k: z[k] for k in z.files})
    one, two = outs
    assert np.allclose(one["split"], two["split"], rtol=2e-5)
    assert np.allclose(one["losses"], two["losses"], rtol=2e-5)
    assert np.abs(one["theta"] - two["theta"]).max() <= 2e-5 * np.abs(one["theta"]).max()
