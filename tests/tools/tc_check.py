"""Test tooling (uses the oracle; not product code). Dev check (GPU box): tensor-core class vs the FP64 oracle, per loss component and gradient tensor."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine, rel_inf, layer_slices


def check(dim, inpDim, lw, act, nb, integNum, nbi, bDof, src=False, iw=False, dvec=False, td=True):
    rng = np.random.RandomState(11)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=7) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td, lossOpt=dict(isSource=src, integWflag=iw))
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = make_engine(feed, theta=theta, **kw)
    print(eng.kernel_info())
    try:
        fwd = eng.loss(lossVec=True)
        print("  fwd ", {k: (float(fwd[k]), ref[k]) for k in ("loss", "BCloss", "ICloss", "varLoss")})
        print("  lossVec rel", rel_inf(fwd["lossVec"], ref["lossVec"]))
        out = eng.loss_grad()
        print("  grad", {k: abs(float(out[k]) - ref[k]) / (abs(ref[k]) + 1e-30) for k in ("loss", "BCloss", "ICloss", "varLoss")})
        for name, sl in layer_slices(inpDim, lw):
            print("   %-8s rel %.3e" % (name, rel_inf(out["grad"][sl], ref["grad"][sl])))
        X = rng.uniform(-1, 1, (300, inpDim))
        u = eng.eval(X)
        refu = go.mlp_value(theta.astype(np.float64), X.astype(np.float32).astype(np.float64), inpDim, lw, go.act_id(act))
        print("  eval rel", rel_inf(u, refu))
    finally:
        eng.close()


if __name__ == "__main__":
    check(2, 3, [128, 128], "tanh", 40, 64, 300, 200)
    check(1, 2, [100, 256, 130], "sigmoid", 70, 16, 150, 100, src=True)
    check(2, 3, [256, 256, 256, 256], "tanh", 700, 64, 3000, 2000)
    check(2, 5, [80], "tanh", 33, 36, 70, 40, iw=True, dvec=True)
