"""GPU parity of the thread-per-point kernel for narrow networks (vn_tpp.cu: every hidden width <= 32 — the reference's own
operator configurations) against the FP64 oracle, through the C ABI.  Same bar as the other kernel classes: 1e-5 relative for
the loss, its components, lossVec and every gradient tensor; the achieved errors are printed (`pytest -s`)."""
import numpy as np
import pytest

from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine, rel_inf, layer_slices

TOL = 1e-5
pytestmark = pytest.mark.gpu

CASES = [
    # dim inpDim layers        act        td     src    iw     dvec   nb    integNum nbi  bDof
    (1, 2, [20], "sigmoid", True, False, False, False, 6000, 16, 640, 600),          # Operator_1Dt shape (96 000 points)
    (2, 3, [10, 20], "sigmoid", True, False, False, False, 2100, 64, 333, 200),      # Operator_2Dt network, 1 050 tiles
    (1, 3, [10, 20, 30], "sigmoid", True, False, False, False, 6000, 16, 1750, 1600),  # Operator_1DtMOR mini-batch shape
    (2, 3, [10, 20], "tanh", True, True, True, True, 333, 32, 70, 40),               # source, Gauss weights, detJ vector
    (2, 2, [32, 7], "tanh", False, True, False, False, 53, 128, 60, 60),             # steady 2D, one test function per tile, ragged widths
    (1, 1, [5], "sigmoid", False, False, False, False, 999, 8, 40, 40),              # steady 1D, tiny network, 16 test functions per tile
    (2, 5, [24, 16, 8], "tanh", True, False, False, False, 129, 64, 70, 40),         # extra (MOR-like) inputs in the table, ragged tile count
    (1, 2, [32, 32], "sigmoid", True, False, True, False, 77, 4, 50, 30),            # widest supported layers, integNum 4
]


def _ids(c):
    return "%s%s_d%d_q%d" % (c[2], c[3], c[0], c[9])


@pytest.mark.parametrize("case", CASES, ids=[_ids(c) for c in CASES])
def test_thread_per_point_kernel_matches_oracle(case):
    dim, inpDim, lw, act, td, src, iw, dvec, nb, integNum, nbi, bDof = case
    rng = np.random.RandomState(977 + nb)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=7) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td, lossOpt=dict(isSource=src, integWflag=iw))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        assert "family=fp32-thread-per-point" in eng.kernel_info(), eng.kernel_info()
        out = eng.loss_grad()
        lv_kernel = eng.get_lossvec()                        # R_i / lossVec as written by the adjoint launch itself
        ach = {}
        for k in ("loss", "BCloss", "ICloss", "varLoss"):
            ach[k] = abs(float(out[k]) - ref[k]) / max(abs(ref[k]), 1e-300)
            assert abs(float(out[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, out[k], ref[k])
        gmax = 0.0
        for name, sl in layer_slices(inpDim, lw)[:-1]:
            e = rel_inf(out["grad"][sl], ref["grad"][sl])
            gmax = max(gmax, e)
            assert e <= TOL, (name, e)
        assert abs(float(out["grad"][-1]) - float(ref["grad"][-1])) <= go.bout_tolerance(ref, feed, td), "output bias"
        ach["grad"] = gmax
        ach["lossVec"] = rel_inf(lv_kernel, ref["lossVec"])
        assert ach["lossVec"] <= TOL
        again = eng.loss_grad()                              # fixed-order reductions: bitwise reproducible
        assert np.array_equal(again["grad"], out["grad"]) and float(again["loss"]) == float(out["loss"])
        print("[achieved] tpp %-28s %s" % (_ids(case), "  ".join("%s=%.2e" % kv for kv in ach.items())))
    finally:
        eng.close()


def test_thread_per_point_agrees_with_fma_tile_class(monkeypatch):
    """Both kernel families on the same feed (VARNET_B200_CLASS=fma keeps narrow networks on the FMA tiles): many tiles per
    CTA, a ragged last tile."""
    rng = np.random.RandomState(5)
    dim, inpDim, lw = 2, 3, [10, 20]
    nb = 148 * 2 * 23 + 5
    feed = synth_feed(rng, dim, inpDim, nb, 64, 333, 200)
    theta = go.glorot_init(inpDim, lw, seed=3)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    outs = {}
    for kind in ("fma", "tpp"):
        if kind == "fma":
            monkeypatch.setenv("VARNET_B200_CLASS", "fma")
        else:
            monkeypatch.delenv("VARNET_B200_CLASS", raising=False)
        eng = make_engine(feed, theta=theta, **kw)
        try:
            assert ("thread-per-point" in eng.kernel_info()) == (kind == "tpp"), eng.kernel_info()
            outs[kind] = eng.loss_grad()
        finally:
            eng.close()
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(outs["tpp"][k]) - float(outs["fma"][k])) <= TOL * abs(float(outs["fma"][k]))
    for name, sl in layer_slices(inpDim, lw)[:-1]:
        assert rel_inf(outs["tpp"]["grad"][sl], outs["fma"]["grad"][sl]) <= TOL, name


def test_thread_per_point_training_minibatch_and_extra_inputs():
    """Adam steps through vn_train_step follow the oracle's TF-Adam trajectory; a device-resident index list (vn_set_batch)
    selects test functions like a host-side gather (the MOR network of Operator_1DtMOR; per-call constant inputs are covered by
    tests/test_gpu_configs.py, which runs the narrow operator configurations through this kernel)."""
    rng = np.random.RandomState(9)
    dim, inpDim, lw = 1, 3, [10, 20, 30]
    nb, integNum = 96, 16
    feed = synth_feed(rng, dim, inpDim, nb, integNum, 120, 80)
    theta = go.glorot_init(inpDim, lw, seed=11)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    eng = make_engine(feed, theta=theta, **kw)
    try:
        assert "thread-per-point" in eng.kernel_info()
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


def test_train_batches_equals_set_batch_plus_train_step():
    """vn_train_batches (k mini-batches of the resident table, one host round trip) takes bit for bit the steps of
    k x {vn_set_batch, vn_train_step}: same losses, same weights, the engine left on the last batch."""
    rng = np.random.RandomState(21)
    dim, inpDim, lw = 1, 3, [10, 20, 30]
    nb, integNum, k, nbatch = 240, 16, 6, 40
    feed = synth_feed(rng, dim, inpDim, nb, integNum, 120, 80)
    theta = go.glorot_init(inpDim, lw, seed=11)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    idx = np.stack([rng.permutation(nb)[:nbatch] for _ in range(k)]).astype(np.int32)
    res = []
    for batched in (0, 1, 2):
        eng = make_engine(feed, theta=theta, **kw)
        try:
            if batched == 1:
                losses = eng.train_batches(1e-3, idx)
            elif batched == 2:
                # the two halves (vn_train_batches_begin / _end), two calls in flight
                eng.train_batches_begin(1e-3, idx[:k // 3])
                eng.train_batches_begin(1e-3, idx[k // 3:2 * (k // 3)])
                with pytest.raises(Exception):
                    eng.train_batches_begin(1e-3, idx[:1])               # a third call in flight is refused
                with pytest.raises(Exception):
                    eng.train_batches(1e-3, idx[:1])                     # and so is the waiting form
                first = eng.train_batches_end()                          # the older one
                eng.train_batches_begin(1e-3, idx[2 * (k // 3):])
                losses = np.concatenate([first, eng.train_batches_end(), eng.train_batches_end()])
                with pytest.raises(Exception):
                    eng.train_batches_end()
            else:
                losses = []
                for row in idx:
                    eng.set_batch(row)
                    losses.append(eng.train_step(1e-3))
            last = eng.loss_grad()                       # on the last batch, with the final weights
            res.append((np.asarray(losses, dtype=np.float32), eng.get_params().copy(), float(last["loss"])))
        finally:
            eng.close()
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) and res[0][2] == r[2]
