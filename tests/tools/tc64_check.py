"""Development probe for the width-64 tensor-core tile kernel (vn_tc64.cu): per-tensor errors against the FP64
oracle and against the FMA class, plus a timing comparison on a larger synthetic batch.  Not a test."""
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine, rel_inf, layer_slices

CASES = [
    (2, 3, [64, 64, 64, 64], "tanh", True, False, False, False, 300, 64, 333, 200),
    (1, 2, [64, 64], "sigmoid", True, True, False, False, 70, 16, 150, 100),
    (2, 3, [40, 64, 50], "tanh", True, True, True, True, 33, 32, 70, 40),
    (2, 5, [48, 33, 64, 64, 40], "sigmoid", True, False, False, False, 129, 64, 70, 40),
    (1, 2, [64, 64, 64, 64, 64, 64], "tanh", True, False, False, False, 300, 32, 70, 40),
]


def run_case(case):
    dim, inpDim, lw, act, td, src, iw, dvec, nb, integNum, nbi, bDof = case
    rng = np.random.RandomState(4321 + nb)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=7) + 0.05 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td, lossOpt=dict(isSource=src, integWflag=iw))
    ref = go.loss_and_grad(theta, feed, **kw)
    outs = {}
    for kind in ("fma", "tc64"):
        os.environ["VARNET_B200_CLASS"] = kind
        eng = make_engine(feed, theta=theta, **kw)
        try:
            info = eng.kernel_info().split()[0]
            outs[kind] = eng.loss_grad()
            outs[kind + "_R"] = None
        except Exception as ex:
            print("  %s FAILED: %r" % (kind, ex))
            outs[kind] = None
        finally:
            eng.close()
        print("  %s -> %s" % (kind, info))
    for kind in ("fma", "tc64"):
        o = outs[kind]
        if o is None:
            continue
        print("  [%s] loss %.3e varLoss %.3e" % (kind, abs(float(o["loss"]) - ref["loss"]) / abs(ref["loss"]),
                                                 abs(float(o["varLoss"]) - ref["varLoss"]) / abs(ref["varLoss"])))
        for name, sl in layer_slices(inpDim, lw):
            print("      %-10s rel_inf %.3e" % (name, rel_inf(o["grad"][sl], ref["grad"][sl])))


def timing():
    rng = np.random.RandomState(1)
    dim, inpDim, lw = 2, 3, [64, 64, 64, 64]
    nb = 148 * 2 * 64 * 2
    feed = synth_feed(rng, dim, inpDim, nb, 64, 1000, 600)
    theta = go.glorot_init(inpDim, lw, seed=3)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    res = {}
    for kind in ("fma", "tc64"):
        os.environ["VARNET_B200_CLASS"] = kind
        eng = make_engine(feed, theta=theta, **kw)
        try:
            for _ in range(2):
                o = eng.loss_grad()
            t0 = time.perf_counter()
            for _ in range(5):
                o = eng.loss_grad()
            dt = (time.perf_counter() - t0) / 5
            res[kind] = o
            print("  timing %s: %.3f ms per loss_grad, %.1f M pts/s  loss %.6e" % (kind, dt * 1e3, nb * 64 / dt / 1e6, float(o["loss"])))
        finally:
            eng.close()
    if len(res) == 2:
        for name, sl in layer_slices(inpDim, lw):
            print("      %-10s tc64 vs fma rel_inf %.3e" % (name, rel_inf(res["tc64"]["grad"][sl], res["fma"]["grad"][sl])))
        o1 = res["tc64"]
    # determinism
    os.environ["VARNET_B200_CLASS"] = "tc64"
    eng = make_engine(feed, theta=theta, **kw)
    try:
        a = eng.loss_grad()["grad"].copy(); b = eng.loss_grad()["grad"].copy()
        print("  bitwise reproducible:", bool(np.array_equal(a, b)))
    finally:
        eng.close()


if __name__ == "__main__":
    for c in CASES:
        print("case", c[:4], "nb", c[8], "integNum", c[9])
        run_case(c)
    if "--timing" in sys.argv:
        timing()
