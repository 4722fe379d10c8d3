"""Worker for the world-size-2 tests: every rank builds the same tiny Operator_1Dt problem through the
host mirror with processors=['GPU:0','GPU:1'], owns the tower whose index equals its rank, takes
`steps` optimizer steps through TFNN.sess.run and writes loss history + final weights to an .npz.

  CPU (gloo, tests/fake_engine.py injected):  python tests/mp_worker.py cpu <out> <steps>   (spawned)
  GPU (nccl, real engine, under torchrun):    torchrun --nproc-per-node 2 tests/mp_worker.py gpu <out> <steps>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(mode, out, steps, world, rank, batchNum=None, scale=0.2):
    import torch
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        if mode == "gpu":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    import varnet_b200
    import varnet_b200.backend as be
    if mode == "cpu":
        from tests.fake_engine import FakeEngine
        be.Engine = FakeEngine
    from oracle import configs
    api = varnet_b200
    dom = api.Domain1D()
    pde = api.ADPDE(dom, diff=0.1 / np.pi, vel=1.0, timeDependent=True, tInterval=[0, 2.0], IC=lambda x: -np.sin(np.pi * x))
    procs = ['GPU:%d' % i for i in range(world)] if world > 1 else 'GPU:0'
    vn = api.VarNet(pde, layerWidth=[12], discNum=max(2, int(20 * scale)), bDiscNum=None, tDiscNum=max(2, int(300 * scale)),
                    processors=procs, seed=123)
    tf = vn.tfData
    fd = vn.fixData
    fd.setFEdata()
    Input, _, biInput, _ = vn.trainingPoints()
    tData = api.ManageTrainData(Input, biInput, batchNum, None, False, 1)
    tData = vn.trainData(0, None, tData)
    tData.updateDictFields('trainW', np.array([10., 10., 1.]))       # divides the BC/IC weights by batchNum*puNum in place
    bc, ic, var, lv = tData.splitLoss(tf, True)
    losses = [float(tData.optimIter(tf)) for _ in range(steps)]
    theta = tf.get_parameters()
    if rank == 0:
        np.savez(out, losses=np.array(losses), theta=theta, split=np.array([bc, ic, var], dtype=np.float64), lossVec=lv)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _spawned(rank, world, out, steps, batchNum):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=os.environ.get("MASTER_PORT", "29533"))
    run("cpu", out, steps, world, rank, batchNum)


if __name__ == "__main__":
    mode, out, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    run(mode, out, steps, world, rank)
