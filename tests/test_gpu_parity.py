"""GPU parity: CUDA engine (through the C ABI) vs the FP64 oracle on identical seeded feeds.

Tolerance (BASELINE.json north_star): 1e-5 relative in FP32 for the loss, each loss component, the
lossVec field (max-norm relative) and each gradient tensor (||d||inf / ||g||inf)."""
import numpy as np
import pytest

from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine, rel_inf, layer_slices

TOL = 1e-5

CASES = [
    # dim inpDim layers          act       td     src    iw     dvec   nb   integNum nbi bDof
    (1, 2, [20], "sigmoid", True, False, False, False, 300, 16, 62, 40),            # Operator_1Dt shape
    (2, 3, [10, 20], "sigmoid", True, False, False, False, 150, 64, 212, 180),      # Operator_2Dt shape
    (1, 3, [10, 20, 30], "sigmoid", True, False, False, False, 250, 16, 175, 160),  # Operator_1DtMOR shape
    (2, 3, [64, 64, 64, 64], "tanh", True, False, False, False, 100, 64, 300, 200), # synthetic 2D+t scale-up net
    (2, 3, [16, 16, 16, 16], "tanh", True, True, False, False, 37, 64, 100, 77),    # width sweep 16 + source
    (1, 2, [20], "tanh", True, True, True, False, 41, 36, 50, 33),                  # integPnum=3 (integNum=36, weights)
    (2, 2, [12, 7], "sigmoid", False, True, False, False, 53, 16, 60, 60),          # steady 2D, no IC rows
    (2, 3, [24, 24], "tanh", True, False, True, True, 19, 216, 90, 50),             # integPnum=3 2D+t, vector detJ
    (1, 1, [9], "sigmoid", False, False, False, False, 77, 4, 2, 2),                # steady 1D, tiny
    (2, 5, [33, 40, 64], "sigmoid", True, True, False, True, 29, 64, 129, 100),     # MOR-like extra inputs, ragged widths
    # narrow class (width <= 16, 256-point tiles)
    (2, 3, [16, 9, 16], "sigmoid", True, False, True, True, 61, 36, 70, 30),            # two-pass (36), vector detJ
    (1, 2, [8], "tanh", True, True, False, False, 700, 16, 300, 200),                   # several 256-point tiles
    # every hidden layer exactly 64 wide
    (1, 2, [64, 64], "sigmoid", True, True, True, False, 83, 16, 70, 41),           # 1D+t, fused single pass
    (2, 3, [64, 64, 64], "tanh", True, True, True, True, 23, 36, 45, 30),           # two-pass (36 does not divide the tile)
    (2, 4, [64], "tanh", True, False, False, False, 31, 64, 40, 25),                # one hidden layer + extra input
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c[2]) + c[3] + ("_d%d" % c[0]) for c in CASES])
def test_loss_and_grad_match_oracle(case):
    dim, inpDim, lw, act, td, src, iw, dvec, nb, integNum, nbi, bDof = case
    rng = np.random.RandomState(1234 + nb)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=7) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    lossOpt = dict(isSource=src, integWflag=iw)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td, lossOpt=lossOpt)
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        fwd = eng.loss(lossVec=True)
        for k in ("loss", "BCloss", "ICloss", "varLoss"):
            assert abs(float(fwd[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, fwd[k], ref[k])
        assert rel_inf(fwd["lossVec"], ref["lossVec"]) <= TOL
        out = eng.loss_grad()
        for k in ("loss", "BCloss", "ICloss", "varLoss"):
            assert abs(float(out[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, out[k], ref[k])
        for name, sl in layer_slices(inpDim, lw):
            assert rel_inf(out["grad"][sl], ref["grad"][sl]) <= TOL, name
        assert eng.launch_count() > 0
    finally:
        eng.close()


@pytest.mark.gpu
def test_eval_matches_oracle():
    rng = np.random.RandomState(5)
    inpDim, lw = 3, [64, 64, 64, 64]
    theta = go.glorot_init(inpDim, lw, seed=1)
    from varnet_b200._capi import Engine
    eng = Engine(2, inpDim, lw, "tanh", True)
    eng.set_params(theta)
    X = rng.uniform(-1, 1, (1000, inpDim))
    u = eng.eval(X)
    ref = go.mlp_value(theta.astype(np.float64), X.astype(np.float32).astype(np.float64), inpDim, lw, go.ACT_TANH)
    assert rel_inf(u, ref) <= TOL
    eng.close()


@pytest.mark.gpu
def test_f32_and_f64_uploads_agree_bitwise():
    rng = np.random.RandomState(11)
    feed = synth_feed(rng, 1, 2, 64, 16, 40, 30)
    lw = [20]
    theta = go.glorot_init(2, lw, seed=2)
    kw = dict(dim=1, inpDim=2, layerWidth=lw, activation="sigmoid", timeDependent=True,
              lossOpt=dict(isSource=False, integWflag=False))
    e64 = make_engine(feed, theta=theta, dtype=np.float64, **kw)
    e32 = make_engine(feed, theta=theta, dtype=np.float32, **kw)
    a, b = e64.loss_grad(), e32.loss_grad()
    assert np.array_equal(a["grad"], b["grad"]) and a["loss"] == b["loss"]
    e64.close(); e32.close()


@pytest.mark.gpu
def test_adam_matches_tf_formula():
    rng = np.random.RandomState(3)
    feed = synth_feed(rng, 1, 2, 128, 16, 40, 30)
    lw = [20]
    theta = go.glorot_init(2, lw, seed=4)
    kw = dict(dim=1, inpDim=2, layerWidth=lw, activation="sigmoid", timeDependent=True,
              lossOpt=dict(isSource=False, integWflag=False))
    eng = make_engine(feed, theta=theta, **kw)
    th = theta.astype(np.float64); m = np.zeros_like(th); v = np.zeros_like(th)
    for t in range(1, 6):
        ref = go.loss_and_grad(th.astype(np.float32), feed, **kw)
        loss = eng.train_step(1e-3)
        assert abs(float(loss) - ref["loss"]) <= 2e-5 * abs(ref["loss"])
        th, m, v = go.adam_step(th, ref["grad"], m, v, t, lr=1e-3)
        assert rel_inf(eng.get_params(), th) <= 2e-5
    eng.close()


RES_CASES = [(1, 2, [20], "sigmoid", True), (2, 3, [10, 20], "sigmoid", True), (2, 3, [64, 64, 64, 64], "tanh", True),
             (2, 2, [16, 24], "tanh", False), (1, 3, [10, 20, 30], "sigmoid", True), (1, 1, [12], "tanh", False),
             # wide networks (tensor-core class): plain FP32 residual kernel of vn_tc.cu
             (2, 3, [256, 256], "tanh", True), (1, 4, [100, 130, 72], "tanh", True), (2, 2, [128], "tanh", False)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", RES_CASES, ids=[str(c[2]) + c[3] + ("_d%d" % c[0]) for c in RES_CASES])
def test_strong_residual_matches_oracle(case):
    """vn_residual_f64 vs NNModel.Residual restated (TFModel.py:718-772): u_t, Laplacian by
    forward-over-forward second derivatives."""
    dim, inpDim, lw, act, td = case
    rng = np.random.RandomState(77)
    n = 777
    theta = go.glorot_init(inpDim, lw, seed=9) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    X = rng.uniform(-1, 1, (n, inpDim)); diff = rng.rand(n, 1) * 0.1; vel = rng.randn(n, dim)
    ddx = rng.randn(n, dim) * 0.01; src = rng.randn(n, 1)
    from varnet_b200._capi import Engine
    eng = Engine(dim, inpDim, lw, act, td)
    eng.set_params(theta)
    u, res = eng.residual(X, diff, vel, ddx, src)
    uo, ro = go.strong_residual(theta, X, diff, vel, ddx, src, dim, inpDim, lw, act, td)
    assert rel_inf(u, uo) <= TOL
    # the residual mixes terms of different size: bound relative to the sum of their magnitudes
    assert np.abs(res - ro).max() <= TOL * np.abs(ro).max() + 1e-6
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dvec", [False, True])
def test_indexed_batch_and_extra_inputs_equal_materialised_upload(dvec):
    """Device-resident mini-batch (table + test-function index list + constant MOR input) must be bitwise
    identical to uploading the gathered rows the way the reference feeds them (VarNetUtility.py:833-844)."""
    from varnet_b200._capi import Engine
    rng = np.random.RandomState(21)
    dim, inpDim, lw, nbTab, q = 1, 3, [10, 20, 30], 50, 16
    full = synth_feed(rng, dim, 2, nbTab, q, 40, 30, True, False, False, dvec)
    mor = 0.0123
    theta = go.glorot_init(inpDim, lw, seed=3)
    perm = rng.permutation(nbTab)[:23]
    rows = (perm.reshape(-1, 1) * q + np.arange(q)).ravel()
    bX = np.hstack([full["biInput"], np.full((40, 1), mor)])

    a = Engine(dim, inpDim, lw, "sigmoid", True)
    a.set_params(theta)
    a.select_table(3)
    a.upload_table(full["Input"], full["gcoef"], None, None, full["dNt"], [nbTab, q], None, full["detJ"], dvec, nx=2)
    a.set_extra_inputs([mor])
    a.set_batch(perm)
    a.upload_bic(bX, full["biLabel"], 30, 2.0)
    a.set_weights(full["w"])

    b = Engine(dim, inpDim, lw, "sigmoid", True)
    b.set_params(theta)
    Xm = np.hstack([full["Input"][rows], np.full((len(rows), 1), mor)])
    detJ = full["detJ"][perm] if dvec else full["detJ"]
    b.upload_points(Xm, full["gcoef"][rows], None, None, full["dNt"][rows], [len(perm), q], None, detJ, dvec)
    b.upload_bic(bX, full["biLabel"], 30, 2.0)
    b.set_weights(full["w"])

    ra, rb = a.loss_grad(), b.loss_grad()
    assert ra["loss"] == rb["loss"] and np.array_equal(ra["grad"], rb["grad"])
    la, lb = a.loss(lossVec=True), b.loss(lossVec=True)
    assert np.array_equal(la["lossVec"], lb["lossVec"]) and la["varLoss"] == lb["varLoss"]
    # the same table serves another batch without any re-upload; training steps replay a captured graph
    a.set_batch(np.arange(nbTab))
    l1 = a.train_step(1e-3); l2 = a.train_step(1e-3)
    assert np.isfinite(l1) and np.isfinite(l2)
    a.close(); b.close()


@pytest.mark.gpu
def test_c_abi_error_behaviour():
    """Error conventions of the boundary: int status + message, mirroring the reference's ValueError /
    'trainDicts must be called first' sites (TFModel.py:117-134, VarNetUtility.py:1031-1032)."""
    from varnet_b200._capi import Engine, EngineError
    eng = Engine(1, 2, [8], "sigmoid", True)
    with pytest.raises(EngineError, match="must be called first"):
        eng.loss_grad()
    with pytest.raises(EngineError, match="must be called first"):
        eng.train_step(1e-3)
    with pytest.raises(EngineError, match="parameter count mismatch"):
        eng.set_params(np.zeros(3, dtype=np.float32))
    rng = np.random.RandomState(0)
    feed = synth_feed(rng, 1, 2, 8, 6, 10, 6)                        # integNum = 6: not a multiple of 4
    eng.set_params(go.glorot_init(2, [8], seed=0))
    eng.upload_points(feed["Input"], feed["gcoef"], None, None, feed["dNt"], feed["intShape"], None, feed["detJ"])
    with pytest.raises(EngineError, match="must be called first"):   # BC/IC rows still missing
        eng.loss()
    eng.upload_bic(feed["biInput"], feed["biLabel"], 6, 2.0)
    with pytest.raises(EngineError, match="multiple of 4"):
        eng.set_batch(np.arange(4))
    with pytest.raises(EngineError, match="learning rate must be positive"):
        eng.train_step(-1.0)
    with pytest.raises(EngineError, match="dNt is required"):
        eng.upload_points(feed["Input"], feed["gcoef"], None, None, None, feed["intShape"], None, feed["detJ"])
    out = eng.loss_grad()                                             # engine still usable after the errors
    ref = go.loss_and_grad(go.glorot_init(2, [8], seed=0), dict(feed, w=np.ones(3)), dim=1, inpDim=2, layerWidth=[8],
                           activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    assert abs(float(out["loss"]) - ref["loss"]) <= TOL * abs(ref["loss"])
    with pytest.raises(EngineError, match="exceeds the compiled kernel families"):
        Engine(2, 3, [300, 300], "tanh", True)
    deep = Engine(2, 3, [64] * 7, "tanh", True)                       # deep 64-wide nets use the 32-point-tile class
    assert "class=164" in deep.kernel_info()
    deep.close(); eng.close()


@pytest.mark.gpu
def test_deep_and_ragged_networks_match_oracle():
    rng = np.random.RandomState(31)
    for dim, inpDim, lw, act in ((2, 3, [64] * 6, "tanh"), (1, 2, [5, 64, 3, 17, 40], "sigmoid")):
        feed = synth_feed(rng, dim, inpDim, 45, 16, 33, 20)
        theta = go.glorot_init(inpDim, lw, seed=5)
        kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=True,
                  lossOpt=dict(isSource=False, integWflag=False))
        ref = go.loss_and_grad(theta, feed, **kw)
        eng = make_engine(feed, theta=theta, **kw)
        out = eng.loss_grad()
        assert abs(float(out["loss"]) - ref["loss"]) <= TOL * abs(ref["loss"])
        for name, sl in layer_slices(inpDim, lw):
            assert rel_inf(out["grad"][sl], ref["grad"][sl]) <= TOL, name
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lw,nb,q", [([16, 16], 125000, 36), ([128, 128], 70000, 64)], ids=["two_pass_q36", "tensor_core_class_w128"])
def test_fed_step_other_kernel_families(lw, nb, q):
    """vn_loss_grad_fed on the kernel families that used to fall back to upload-then-step: the two-pass class (integNum = 36 divides
    no tile: the forward pass now runs per uploaded chunk) and the tensor-core class (its point chunks wait for the uploads that
    cover them).  Same loss / lossVec bits and gradient as vn_upload_points + vn_loss_grad on a table of more than one 4 Mi-row chunk."""
    from varnet_b200._capi import Engine
    import ctypes as C
    rng = np.random.RandomState(13)
    dim, inpDim = 2, 3
    P = nb * q
    X = rng.uniform(-1, 1, (P, inpDim)); G = rng.randn(P, dim); dNt = rng.randn(P, 1)
    bX = rng.uniform(-1, 1, (500, inpDim)).astype(np.float32); bL = rng.randn(500, 1).astype(np.float32)
    theta = go.glorot_init(inpDim, lw, seed=2)
    outs = []
    for fed in (False, True):
        eng = Engine(dim, inpDim, lw, "tanh", True)
        try:
            eng.set_params(theta)
            eng.upload_bic(bX, bL, 300, 2.0)
            eng.set_weights([3.0, 5.0, 7.0])
            if fed:
                eng.loss_grad_fed(X, G, None, None, dNt, [nb, q], None, 1.3e-6, False)
                g = np.empty(eng.nparam, dtype=np.float32); o = np.empty(4, dtype=np.float32)
                eng._check(eng.lib.vn_get_grad(eng._h, g.ctypes.data_as(C.POINTER(C.c_float)), g.size, o.ctypes.data_as(C.POINTER(C.c_float))))
                r = dict(loss=o[0], varLoss=o[3], grad=g)
            else:
                eng.upload_points(X, G, None, None, dNt, [nb, q], None, 1.3e-6, False)
                r = eng.loss_grad()
            outs.append((r, eng.get_lossvec().copy(), eng.kernel_info()))
        finally:
            eng.close()
    a, b = outs
    assert ("tcgen05" in a[2]) == (lw[0] > 64) and ("two-pass" in a[2]) == (q == 36), a[2]
    for k in ("loss", "varLoss"):
        assert abs(float(a[0][k]) - float(b[0][k])) <= 1e-6 * abs(float(a[0][k])), k
    assert rel_inf(b[0]["grad"], a[0]["grad"]) <= 1e-6
    assert np.array_equal(a[1], b[1])


@pytest.mark.gpu
@pytest.mark.parametrize("feed_dtype", [np.float32, np.float64])
def test_fed_step_overlapping_copies_equals_upload_then_step(feed_dtype):
    """vn_loss_grad_fed (table uploaded chunk by chunk while the adjoint kernel already runs on the first chunks; pageable NumPy
    arrays — float32 or the reference's float64 — go through the multi-threaded pinned bounce buffers)
    against vn_upload_points + vn_loss_grad on the same feed, on a table of more than one 4 Mi-row chunk; then the
    shim path: sess.run with feed_cache=False goes through the fed call and follows the same Adam trajectory."""
    from varnet_b200._capi import Engine
    rng = np.random.RandomState(12)
    dim, inpDim, lw, nb, q = 2, 3, [16, 16], 70000, 64                       # 4.48 M points = 2 chunks
    P = nb * q
    X = rng.uniform(-1, 1, (P, inpDim)).astype(np.float32); G = rng.randn(P, dim).astype(np.float32)
    dNt = rng.randn(P, 1).astype(np.float32)
    bX = rng.uniform(-1, 1, (500, inpDim)).astype(np.float32); bL = rng.randn(500, 1).astype(np.float32)
    theta = go.glorot_init(inpDim, lw, seed=2)
    outs = []
    for fed in (False, True):
        eng = Engine(dim, inpDim, lw, "tanh", True)
        eng.set_params(theta)
        eng.upload_bic(bX, bL, 300, 2.0)
        eng.set_weights([3.0, 5.0, 7.0])
        if fed:
            eng.loss_grad_fed(X.astype(feed_dtype), G.astype(feed_dtype), None, None, dNt.astype(feed_dtype), [nb, q], None, 1.3e-6, False)
        else:
            eng.upload_points(X, G, None, None, dNt, [nb, q], None, 1.3e-6, False)
        r = eng.loss_grad() if not fed else None
        if fed:
            g = np.empty(eng.nparam, dtype=np.float32); o = np.empty(4, dtype=np.float32)
            import ctypes as C
            eng._check(eng.lib.vn_get_grad(eng._h, g.ctypes.data_as(C.POINTER(C.c_float)), g.size, o.ctypes.data_as(C.POINTER(C.c_float))))
            r = dict(loss=o[0], BCloss=o[1], ICloss=o[2], varLoss=o[3], grad=g)
            lv = eng.loss(lossVec=True)["lossVec"]
        else:
            lv = eng.loss(lossVec=True)["lossVec"]
        outs.append((r, lv))
        eng.close()
    a, b = outs
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(a[0][k]) - float(b[0][k])) <= 1e-6 * abs(float(a[0][k])), k
    assert rel_inf(b[0]["grad"], a[0]["grad"]) <= 1e-6
    assert np.array_equal(a[1], b[1])


@pytest.mark.gpu
def test_fed_step_from_registered_feed_arrays_is_bit_identical_to_the_staged_one():
    """Feed arrays that show up a second time are page-locked in place (_capi.HostPins -> vn_host_register) and the fed step copies
    them by DMA as float64, cast by the pack kernel, instead of casting them through bounce buffers with host threads: same
    rounding (round-to-nearest-even on both sides), so loss, gradient and lossVec keep their bits; in-place edits of a registered
    array are seen by the next step; release() unregisters."""
    import ctypes as C
    from varnet_b200._capi import Engine, HostPins
    rng = np.random.RandomState(14)
    dim, inpDim, lw, nb, q = 2, 3, [16, 16], 70000, 64                       # 4.48 M points = 2 chunks, 108 MB + 72 MB + 36 MB of float64
    P = nb * q
    X = rng.uniform(-1, 1, (P, inpDim)); G = rng.randn(P, dim); dNt = rng.randn(P, 1)
    bX = rng.uniform(-1, 1, (500, inpDim)); bL = rng.randn(500, 1)
    theta = go.glorot_init(inpDim, lw, seed=2)
    eng = Engine(dim, inpDim, lw, "tanh", True)
    eng.upload_bic(bX, bL, 300, 2.0)
    eng.set_weights([3.0, 5.0, 7.0])

    def step():
        eng.set_params(theta)
        eng.loss_grad_fed(X, G, None, None, dNt, [nb, q], None, 1.3e-6, False)
        g = np.empty(eng.nparam, dtype=np.float32); o = np.empty(4, dtype=np.float32)
        eng._check(eng.lib.vn_get_grad(eng._h, g.ctypes.data_as(C.POINTER(C.c_float)), g.size, o.ctypes.data_as(C.POINTER(C.c_float))))
        return o.copy(), g, eng.get_lossvec()

    pins = HostPins()
    assert not pins.touch_group([X, G, dNt])                                 # first sighting: nothing is registered
    staged = step()
    assert pins.touch_group([X, G, dNt]) and pins.registered == 3            # second sighting: page-locked in place
    assert pins.touch_group([X, G, dNt]) and pins.registered == 3
    direct = step()
    for a, b in zip(staged, direct):
        assert np.array_equal(a, b)
    X[:q] += 0.25                                                            # in-place edit of a registered array
    edited = step()
    assert not np.array_equal(edited[2][:1], direct[2][:1]) and np.array_equal(edited[2][1:], direct[2][1:])      # lossVec of test function 0 only
    eng.synchronize()
    pins.release()
    assert pins.bytes == 0 and not pins.pins
    again = step()                                                           # pageable again: the staged path, same bits
    for a, b in zip(edited, again):
        assert np.array_equal(a, b)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("integPnum,lw", [(2, [12, 20]), (3, [12, 20]), (2, [64, 64, 64, 64]), (2, [48, 64])])
def test_device_generated_table_is_bit_identical_to_the_uploaded_one(integPnum, lw):
    """vn_generate_table_f64 (uniform mesh, constant coefficients, table built on the device from the mesh centres and
    the periodic FE tables) against uploading the host-built arrays of the same test-function range: identical loss
    bits, lossVec and gradient.  For the 33-64-wide class no table is materialised at all: the tensor-core tile kernel
    regenerates every row from the centre of its test function (`in-kernel` in kernel_info), also under a mini-batch
    index list."""
    from varnet_b200 import workloads
    from varnet_b200._capi import Engine
    nx, ny, ntime, n0, n1 = 9, 7, 6, 37, 301
    feed, meta = workloads.shard_feed(nx, ny, ntime, n0, n1, integPnum=integPnum, dtype=np.float64)
    theta = go.glorot_init(3, lw, seed=5)
    res = []
    for dev in (False, True):
        eng = Engine(2, 3, lw, "tanh", True, False, meta["lossOpt"]["integWflag"])
        eng.set_params(theta)
        if dev:
            bic, _ = workloads.generate_on_device(eng, nx, ny, ntime, n0, n1, integPnum=integPnum)
            assert bic["intShape"] == feed["intShape"]
        else:
            eng.upload_points(feed["Input"], feed["gcoef"], feed["source"], feed["N"], feed["dNt"], feed["intShape"],
                              feed["integW"], feed["detJ"], False)
        eng.upload_bic(feed["biInput"], feed["biLabel"], feed["bDof"], feed["biDimVal"])
        eng.set_weights([2.0, 3.0, 5.0])
        out = eng.loss_grad()
        lv = eng.loss(lossVec=True)["lossVec"]
        if dev and max(lw) > 32:
            assert "in-kernel" in eng.kernel_info()
        idx = np.random.RandomState(1).permutation(n1 - n0)[:97].astype(np.int32)      # mini-batch = index list into the table
        eng.set_batch(idx)
        sub = eng.loss_grad()
        res.append((out, lv, sub))
        eng.close()
    (a, la, sa), (b, lb, sb) = res
    assert all(np.float32(a[k]) == np.float32(b[k]) for k in ("loss", "BCloss", "ICloss", "varLoss"))
    assert np.array_equal(la, lb) and np.array_equal(a["grad"], b["grad"])
    assert np.float32(sa["loss"]) == np.float32(sb["loss"]) and np.array_equal(sa["grad"], sb["grad"])
