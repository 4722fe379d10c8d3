"""Per-kernel times of the narrow operator configurations (shapes of Operator_1Dt / Operator_1DtMOR mini-batch / Operator_2Dt):
steps/s of the captured step graph (vn_train_steps) and the event-timed kernels of the profiled (non-graph) step."""
import sys
import time
import numpy as np
sys.path.insert(0, ".")
from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine

SHAPES = [
    # name, dim, inpDim, layers, nb, integNum, nbi, bDof
    ("Operator_1Dt", 1, 2, [20], 6000, 16, 640, 600),
    ("Operator_1DtMOR batch", 1, 3, [10, 20, 30], 6000, 16, 1750, 1600),
    ("Operator_2Dt", 2, 3, [10, 20], 240000, 64, 21281, 18000),
]
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, dim, inpDim, lw, nb, q, nbi, bDof in SHAPES:
    if only and only not in name:
        continue
    rng = np.random.RandomState(1)
    feed = synth_feed(rng, dim, inpDim, nb, q, nbi, bDof)
    theta = go.glorot_init(inpDim, lw, seed=5)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="sigmoid", timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    eng = make_engine(feed, theta=theta, dtype=np.float32, **kw)
    k = 256 if nb * q < 1e6 else 8
    eng.train_steps(1e-3, k)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 1.5:
        eng.train_steps(1e-3, k); n += k
    dt = time.perf_counter() - t0
    eng.profile_enable(True); eng.profile_read()
    for _ in range(20):
        eng.train_step(1e-3)
    pr = eng.profile_read()
    per = {a: round(1e3 * b[0] / max(b[1], 1), 2) for a, b in pr.items() if b[1]}
    print("%-24s %9.0f steps/s  %7.1f us/step  %8.1f M pts/s | profiled kernels (us): %s | %s" %
          (name, n / dt, 1e6 * dt / n, nb * q * n / dt / 1e6, per, eng.kernel_info()[:150] + " ... " + eng.kernel_info()[-6:]), flush=True)
    eng.close()
