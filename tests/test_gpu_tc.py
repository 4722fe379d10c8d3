"""GPU parity of the tensor-core kernel class (hidden width 65..256: tcgen05 3xTF32 layer GEMMs, vn_tc.cu)
against the FP64 oracle, through the C ABI.  Same bar as the FMA classes: 1e-5 relative for the loss, its
components, lossVec and every gradient tensor; the output-bias gradient (a single cancelling sum) uses the
conditioning-aware bound of oracle.graph_oracle.bout_tolerance."""
import numpy as np
import pytest

from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine, rel_inf, layer_slices

TOL = 1e-5

CASES = [
    # dim inpDim layers               act       td     src    iw     dvec   nb   integNum nbi  bDof
    (2, 3, [128, 128], "tanh", True, False, False, False, 40, 64, 300, 200),
    (1, 2, [100, 256, 130], "sigmoid", True, True, False, False, 70, 16, 150, 100),     # ragged widths, 1D+t
    (2, 3, [256, 256, 256, 256], "tanh", True, False, False, False, 700, 64, 3000, 2000),  # width-sweep 256 net
    (2, 5, [80], "tanh", True, False, True, True, 33, 36, 70, 40),                       # one layer, MOR-like inputs, detJ vector
    (2, 2, [96, 72], "sigmoid", False, True, False, False, 53, 16, 60, 60),              # steady 2D, no IC rows
    (1, 3, [65, 200], "tanh", True, True, True, False, 29, 216, 129, 100),               # integNum 216 (chunk = lcm with 128)
]


def check_against_oracle(eng, ref, feed, inpDim, lw, td):
    fwd = eng.loss(lossVec=True)
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(fwd[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, fwd[k], ref[k])
    assert rel_inf(fwd["lossVec"], ref["lossVec"]) <= TOL
    out = eng.loss_grad()
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(out[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, out[k], ref[k])
    slices = layer_slices(inpDim, lw)
    for name, sl in slices[:-1]:
        assert rel_inf(out["grad"][sl], ref["grad"][sl]) <= TOL, name
    assert abs(float(out["grad"][-1]) - float(ref["grad"][-1])) <= go.bout_tolerance(ref, feed, td), "output bias"
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c[2]) + c[3] + ("_d%d" % c[0]) for c in CASES])
def test_tensor_core_class_matches_oracle(case):
    dim, inpDim, lw, act, td, src, iw, dvec, nb, integNum, nbi, bDof = case
    rng = np.random.RandomState(4321 + nb)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=7) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td, lossOpt=dict(isSource=src, integWflag=iw))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        assert "family=tcgen05-3xtf32" in eng.kernel_info()
        check_against_oracle(eng, ref, feed, inpDim, lw, td)
        assert eng.launch_count() > 0
    finally:
        eng.close()


@pytest.mark.gpu
def test_many_chunks_and_eval(monkeypatch):
    """One 128-point tile per SM and chunk: the step runs over several chunks (last one ragged) and the FP64
    accumulation across chunks reproduces the single-pass oracle; vn_eval goes through the same pipeline."""
    monkeypatch.setenv("VARNET_B200_TC_WAVES", "1")
    rng = np.random.RandomState(77)
    dim, inpDim, lw = 2, 3, [128, 96, 128]
    feed = synth_feed(rng, dim, inpDim, 700, 64, 2500, 1700)            # 44800 points > 2 x 18944
    theta = go.glorot_init(inpDim, lw, seed=3)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        assert "chunk=18944" in eng.kernel_info()
        check_against_oracle(eng, ref, feed, inpDim, lw, True)
        X = rng.uniform(-1, 1, (20011, inpDim))
        u = eng.eval(X)
        refu = go.mlp_value(theta.astype(np.float64), X.astype(np.float32).astype(np.float64), inpDim, lw, go.ACT_TANH)
        assert rel_inf(u, refu) <= TOL
    finally:
        eng.close()


@pytest.mark.gpu
def test_forced_tensor_core_class_agrees_with_fma_class(monkeypatch):
    """Width 64 through both kernel families (VARNET_B200_CLASS=tc pads 64 -> 128): the crossover measurement
    of the width sweep compares like with like."""
    rng = np.random.RandomState(5)
    dim, inpDim, lw = 2, 3, [64, 64, 64, 64]
    feed = synth_feed(rng, dim, inpDim, 300, 64, 333, 200)
    theta = go.glorot_init(inpDim, lw, seed=3)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    outs = {}
    for kind in ("fma", "tc"):
        monkeypatch.setenv("VARNET_B200_CLASS", kind)
        eng = make_engine(feed, theta=theta, **kw)
        try:
            assert ("tcgen05" in eng.kernel_info()) == (kind == "tc")
            outs[kind] = eng.loss_grad()
        finally:
            eng.close()
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(float(outs["tc"][k]) - float(outs["fma"][k])) <= TOL * abs(float(outs["fma"][k]))
    for name, sl in layer_slices(inpDim, lw)[:-1]:
        assert rel_inf(outs["tc"]["grad"][sl], outs["fma"]["grad"][sl]) <= TOL, name


@pytest.mark.gpu
def test_training_trajectory_and_minibatch_index_list():
    """Adam steps on the tensor-core class follow the oracle's TF-Adam trajectory; a device-resident index list
    (vn_set_batch) selects test functions of the resident table like a host-side gather."""
    rng = np.random.RandomState(9)
    dim, inpDim, lw = 1, 2, [160, 160]
    nb, integNum = 96, 16
    feed = synth_feed(rng, dim, inpDim, nb, integNum, 120, 80)
    theta = go.glorot_init(inpDim, lw, seed=11)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    eng = make_engine(feed, theta=theta, **kw)
    try:
        th = theta.astype(np.float64); m = np.zeros_like(th); v = np.zeros_like(th)
        for t in range(1, 4):
            ref = go.loss_and_grad(th.astype(np.float32), feed, **kw)
            loss = eng.train_step(1e-3)
            assert abs(float(loss) - ref["loss"]) <= 5e-5 * abs(ref["loss"])
            th, m, v = go.adam_step(th, ref["grad"], m, v, t)
        assert rel_inf(eng.get_params(), th) <= 1e-4
        # mini-batch: 40 of the 96 test functions through the on-device index list
        idx = rng.permutation(nb)[:40]
        eng.set_params(theta)
        eng.set_batch(idx)
        rows = (idx[:, None] * integNum + np.arange(integNum)[None, :]).ravel()
        sub = dict(feed, Input=feed["Input"][rows], gcoef=feed["gcoef"][rows], dNt=feed["dNt"][rows], source=feed["source"][rows],
                   N=feed["N"][rows], intShape=[40, integNum])
        ref = go.loss_and_grad(theta, sub, **kw)
        check_against_oracle(eng, ref, sub, inpDim, lw, True)
    finally:
        eng.close()


@pytest.mark.gpu
def test_tensor_core_class_limits():
    from varnet_b200._capi import Engine, EngineError
    with pytest.raises(EngineError, match="exceeds the compiled kernel families"):
        Engine(2, 3, [300, 300], "tanh", True)
    eng = Engine(2, 3, [256, 256], "tanh", True)
    assert "family=tcgen05-3xtf32" in eng.kernel_info() and "WP=256" in eng.kernel_info()
    with pytest.raises(EngineError, match="must be called first"):
        eng.loss()
    eng.close()


@pytest.mark.gpu
def test_varnet_api_with_wide_network():
    """The reference-facing path (VarNet.train -> TFNN.sess.run, residual monitoring, evaluate) with a 128-wide
    MLP: same host code, the engine picks the tensor-core class by width."""
    import tempfile
    import varnet_b200
    from oracle import configs
    vn = configs.synthetic_2dt(varnet_b200, nx=12, ny=10, ntime=8, layerWidth=(128, 128), activation='tanh', seed=2)
    assert "tcgen05" in vn.tfData.compTowers[0].engine.kernel_info()
    with tempfile.TemporaryDirectory() as d:
        res = vn.train(d, weight=[10., 10., 1.], epochNum=30, saveFreq=10, verbose=False)
        assert len(res.loss) == 30 and np.all(np.isfinite(res.loss))
        assert res.loss[-1] < res.loss[0]
    vn.tfData.sess.close()


@pytest.mark.gpu
def test_shared_memory_operand_pipeline_matches_oracle(monkeypatch):
    """VARNET_B200_TC_TS=0 selects the layer-GEMM variant that keeps both operands in shared memory (three
    rotating accumulator sets); it must meet the same bar as the default A-from-TMEM variant."""
    monkeypatch.setenv("VARNET_B200_TC_TS", "0")
    rng = np.random.RandomState(3)
    dim, inpDim, lw = 2, 3, [200, 256, 96]
    feed = synth_feed(rng, dim, inpDim, 90, 64, 400, 250, True, True)
    theta = go.glorot_init(inpDim, lw, seed=4)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=True, integWflag=False))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        check_against_oracle(eng, ref, feed, inpDim, lw, True)
    finally:
        eng.close()


@pytest.mark.gpu
def test_mn_major_weight_gradient_tiles_match_the_k_major_ones(monkeypatch):
    """VARNET_B200_TC_GW=mn: tc_gw_kernel stores the quad-major 16-byte units as they are into MN-major tiles
    (LayoutType::SWIZZLE_128B_BASE32B) instead of transposing them with scalar stores.  The MMAs contract the same numbers in
    the same order: the result must meet the oracle bar and the weight gradients agree with the default tiles to 1e-6."""
    rng = np.random.RandomState(8)
    dim, inpDim, lw = 2, 3, [200, 256, 96]
    feed = synth_feed(rng, dim, inpDim, 90, 64, 400, 250, True, True)
    theta = go.glorot_init(inpDim, lw, seed=4)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=True, integWflag=False))
    ref = go.loss_and_grad(theta, feed, **kw)
    outs = {}
    for mode in ("k", "mn"):
        monkeypatch.setenv("VARNET_B200_TC_GW", mode)
        eng = make_engine(feed, theta=theta, **kw)
        try:
            outs[mode] = eng.loss_grad()
            check_against_oracle(eng, ref, feed, inpDim, lw, True)
        finally:
            eng.close()
    assert float(outs["mn"]["loss"]) == float(outs["k"]["loss"])
    for name, sl in layer_slices(inpDim, lw)[:-1]:
        # the bias gradients ride on the loaders' FP32 partial sums, whose order follows the loader mapping: the oracle bar
        bar = 1e-6 if name.startswith("kernel") else TOL
        assert rel_inf(outs["mn"]["grad"][sl], outs["k"]["grad"][sl]) <= bar, name


@pytest.mark.gpu
@pytest.mark.parametrize("scale,act", [(4.0, "tanh"), (6.0, "sigmoid"), (0.05, "tanh")])
def test_split_product_is_robust_to_weight_scale(scale, act):
    """Saturated activations (large weights) and tiny pre-activations (small weights): the hi/lo split keeps FP32-level
    parity over the whole exponent range the layers see.  With tiny weights the lower-layer bias gradients are
    cancelling sums of much larger terms (values ~1e-7 against a gradient scale of ~1e-2), so that case bounds the
    error of every tensor by the scale of the whole gradient instead of its own."""
    rng = np.random.RandomState(5)
    dim, inpDim, lw = 2, 3, [256, 192, 256]
    feed = synth_feed(rng, dim, inpDim, 60, 64, 300, 200)
    theta = (go.glorot_init(inpDim, lw, seed=7) * scale).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    try:
        if scale >= 1.0:
            check_against_oracle(eng, ref, feed, inpDim, lw, True)
        else:
            out = eng.loss_grad()
            for k in ("loss", "BCloss", "ICloss", "varLoss"):
                assert abs(float(out[k]) - ref[k]) <= TOL * abs(ref[k]) + 1e-30, (k, out[k], ref[k])
            gscale = np.abs(ref["grad"]).max()
            slices = layer_slices(inpDim, lw)
            for name, sl in slices:
                assert np.abs(out["grad"][sl] - ref["grad"][sl]).max() <= TOL * gscale, name
            for name, sl in slices:                                  # the weight matrices keep the per-tensor bar
                if name.startswith("kernel"):
                    assert rel_inf(out["grad"][sl], ref["grad"][sl]) <= TOL, name
    finally:
        eng.close()
