"""GPU parity of the width-64 tensor-core tile kernel (vn_tc64.cu: 128-point resident tiles, activations in tensor
memory, 3xTF32 tcgen05 layer and weight-gradient GEMMs) against the FP64 oracle, through the C ABI.  Same bar as the
other kernel classes: 1e-5 relative for the loss, its components, lossVec and every gradient tensor."""
import numpy as np
import pytest

from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine, rel_inf, layer_slices

TOL = 1e-5

CASES = [
    # dim inpDim layers               act       td     src    iw     dvec   nb   integNum nbi  bDof
    (2, 3, [64, 64, 64, 64], "tanh", True, False, False, False, 300, 64, 333, 200),      # the headline network
    (1, 2, [64, 64], "sigmoid", True, True, False, False, 70, 16, 150, 100),             # 1D+t, two streams
    (2, 3, [40, 64, 50], "tanh", True, True, True, True, 33, 32, 70, 40),                # ragged widths, detJ vector, Gauss weights
    (2, 5, [48, 33, 64], "sigmoid", True, False, False, False, 129, 64, 70, 40),         # MOR-like extra inputs, ragged tile count
    (2, 2, [64, 36], "tanh", False, True, False, False, 53, 128, 60, 60),                # steady 2D, one test function per tile
    (1, 1, [33, 64, 64], "tanh", False, False, False, False, 999, 8, 40, 40),            # steady 1D, many test functions per tile
    (2, 5, [48, 33, 64, 64, 40], "sigmoid", True, False, False, False, 129, 64, 70, 40), # five hidden layers (deep FMA class for the other kernels)
    (1, 2, [64, 64, 64, 64, 64, 64], "tanh", True, False, False, False, 300, 32, 70, 40),
]


def check_against_oracle(eng, ref, feed, inpDim, lw, td):
    out = eng.loss_grad()
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(out[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, out[k], ref[k])
    slices = layer_slices(inpDim, lw)
    for name, sl in slices[:-1]:
        assert rel_inf(out["grad"][sl], ref["grad"][sl]) <= TOL, name
    assert abs(float(out["grad"][-1]) - float(ref["grad"][-1])) <= go.bout_tolerance(ref, feed, td), "output bias"
    lv = eng.loss(lossVec=True)
    assert rel_inf(lv["lossVec"], ref["lossVec"]) <= TOL
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES[:4], ids=[str(c[2]) + c[3] + ("_d%d" % c[0]) for c in CASES[:4]])
def test_tile64_v2_schedule_matches_oracle(case):
    """The warp-specialised, software-pipelined schedule of the same kernel (VARNET_B200_TC64=v2: issuing warp, mbarrier
    hand-offs, cross-products-first single accumulator) holds the same 1e-5 bar; run in a subprocess because the schedule is
    chosen once per process."""
    import os
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import tests.test_gpu_tc64 as t\n"
            "t.test_tile64_tensor_core_kernel_matches_oracle(%r)\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), case))
    env = dict(os.environ, VARNET_B200_TC64="v2")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c[2]) + c[3] + ("_d%d" % c[0]) for c in CASES])
def test_tile64_tensor_core_kernel_matches_oracle(case):
    dim, inpDim, lw, act, td, src, iw, dvec, nb, integNum, nbi, bDof = case
    rng = np.random.RandomState(4321 + nb)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=7) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td, lossOpt=dict(isSource=src, integWflag=iw))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        assert "family=tcgen05-3xtf32-tile64" in eng.kernel_info()
        check_against_oracle(eng, ref, feed, inpDim, lw, td)
        assert eng.launch_count() > 0
    finally:
        eng.close()


@pytest.mark.gpu
def test_tile64_agrees_with_fma_class_and_is_bitwise_reproducible(monkeypatch):
    """Several tiles per CTA (FP32 window folds, persistent loop), both kernel families on the same feed; two runs of the
    tensor-core kernel give bit-identical gradients (single-writer slab slots, fixed-order reductions)."""
    rng = np.random.RandomState(5)
    dim, inpDim, lw = 2, 3, [64, 64, 64, 64]
    nb = 148 * 2 * 19 + 7                                               # 38 tiles per CTA on 148 SMs, last tile ragged
    feed = synth_feed(rng, dim, inpDim, nb, 64, 333, 200)
    theta = go.glorot_init(inpDim, lw, seed=3)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    outs = {}
    for kind in ("fma", "tc64"):
        monkeypatch.setenv("VARNET_B200_CLASS", kind)
        eng = make_engine(feed, theta=theta, **kw)
        try:
            assert ("tile64" in eng.kernel_info()) == (kind == "tc64")
            outs[kind] = eng.loss_grad()
            if kind == "tc64":
                again = eng.loss_grad()
                assert np.array_equal(again["grad"], outs[kind]["grad"]) and float(again["loss"]) == float(outs[kind]["loss"])
        finally:
            eng.close()
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(outs["tc64"][k]) - float(outs["fma"][k])) <= TOL * abs(float(outs["fma"][k]))
    for name, sl in layer_slices(inpDim, lw)[:-1]:
        assert rel_inf(outs["tc64"]["grad"][sl], outs["fma"]["grad"][sl]) <= TOL, name


@pytest.mark.gpu
def test_tile64_training_trajectory_and_minibatch_index_list():
    """Adam steps through vn_train_step (step graph: weight images re-staged every step) follow the oracle's TF-Adam
    trajectory; a device-resident index list (vn_set_batch) selects test functions like a host-side gather."""
    rng = np.random.RandomState(9)
    dim, inpDim, lw = 1, 2, [64, 48, 64]
    nb, integNum = 96, 16
    feed = synth_feed(rng, dim, inpDim, nb, integNum, 120, 80)
    theta = go.glorot_init(inpDim, lw, seed=11)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    eng = make_engine(feed, theta=theta, **kw)
    try:
        assert "tile64" in eng.kernel_info()
        th = theta.astype(np.float64); m = np.zeros_like(th); v = np.zeros_like(th)
        for t in range(1, 5):
            ref = go.loss_and_grad(th.astype(np.float32), feed, **kw)
            loss = eng.train_step(1e-3)
            assert abs(float(loss) - ref["loss"]) <= 5e-5 * abs(ref["loss"])
            th, m, v = go.adam_step(th, ref["grad"].astype(np.float64), m, v, t, 1e-3)
        assert rel_inf(eng.get_params(), th) <= 1e-4
        idx = rng.permutation(nb)[:40].astype(np.int32)
        sub = dict(feed)
        rows = (idx[:, None] * integNum + np.arange(integNum)[None, :]).ravel()
        for k in ("Input", "gcoef", "source", "N", "dNt"):
            sub[k] = np.asarray(feed[k])[rows]
        sub["intShape"] = [len(idx), integNum]
        cur = eng.get_params()
        ref = go.loss_and_grad(cur, sub, **kw)
        eng.set_batch(idx)
        out = eng.loss_grad()
        assert abs(float(out["loss"]) - ref["loss"]) <= TOL * abs(ref["loss"])
        for name, sl in layer_slices(inpDim, lw)[:-1]:
            assert rel_inf(out["grad"][sl], ref["grad"][sl]) <= TOL, name
    finally:
        eng.close()


@pytest.mark.gpu
def test_mn_major_weight_gradient_operands_are_bit_identical(monkeypatch):
    """VARNET_B200_TC64_GW = 0: every weight-gradient operand written in front of the issue (the default, 3, writes the next step's
    a-operand under the running layer GEMM).  VARNET_B200_TC64_GW = 1 / 2: the weight-gradient GEMM reads MN-major operands (LayoutType::SWIZZLE_128B_BASE32B, the one
    canonical layout in which kind::tf32 takes them: scripts/micro/tc_probe3.cu) written with 16-byte stores instead of the
    transposing scalar stores.  Same MMAs in the same order on the same numbers: loss and gradient must not change by a bit
    (several tiles per CTA, ragged last tile, S = 3 and S = 2)."""
    for dim, inpDim, nb in ((2, 3, 148 * 2 * 3 + 5), (1, 2, 700)):
        rng = np.random.RandomState(17 + dim)
        lw = [64, 48, 64]
        feed = synth_feed(rng, dim, inpDim, nb, 64, 200, 120)
        theta = go.glorot_init(inpDim, lw, seed=3)
        kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
        ref = go.loss_and_grad(theta, feed, **kw)
        outs = []
        for mode in ("3", "0", "1", "2"):
            monkeypatch.setenv("VARNET_B200_TC64_GW", mode)
            eng = make_engine(feed, theta=theta, **kw)
            try:
                assert "tile64" in eng.kernel_info()
                outs.append(eng.loss_grad())
                if mode == "2":
                    check_against_oracle(eng, ref, feed, inpDim, lw, True)
            finally:
                eng.close()
        for o in outs[1:]:
            assert float(o["loss"]) == float(outs[0]["loss"]) and np.array_equal(o["grad"], outs[0]["grad"])
