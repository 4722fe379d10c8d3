"""Test tooling (uses the oracle; not product code). Dev probe (GPU box): times the forward and adjoint kernels on a cfg4-like slice."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import graph_oracle as go
from tests.util import synth_feed, make_engine

def run(nb, lw, act, dim=2, inpDim=3, integNum=64, reps=5):
    rng = np.random.RandomState(0)
    feed = synth_feed(rng, dim, inpDim, nb, integNum, 2048, 1500)
    for k in ("Input", "gcoef", "dNt", "source", "N"):
        feed[k] = np.asarray(feed[k], dtype=np.float32)
    theta = go.glorot_init(inpDim, lw, seed=7)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=True, lossOpt=dict(isSource=False, integWflag=False))
    eng = make_engine(feed, theta=theta, dtype=np.float32, **kw)
    print(eng.kernel_info())
    P = nb * integNum
    for name, fn in (("loss", lambda: eng.lib.vn_loss(eng._h, None, None)), ("loss_grad", lambda: eng.lib.vn_loss_grad(eng._h, None)), ("train_step", lambda: eng.lib.vn_train_step(eng._h, 1e-3, None))):
        fn(); eng.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        eng.synchronize()
        dt = (time.perf_counter() - t0) / reps
        M = sum(i * o for i, o in go.layer_sizes(inpDim, lw))
        print("%-10s lw=%s P=%d  %.3f ms  %.3e pts/s  alg %.2f TFLOP/s" % (name, lw, P, dt * 1e3, P / dt, 6 * (1 + dim) * M * P / dt / 1e12 if name != "loss" else 2 * (1 + dim) * M * P / dt / 1e12))
    eng.profile_enable(True); eng.profile_read()
    for _ in range(reps): eng.lib.vn_train_step(eng._h, 1e-3, None)
    pr = eng.profile_read()
    print("   per-kernel us:", {k: round(v[0] * 1e3 / max(v[1], 1), 1) for k, v in pr.items()})
    eng.close()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if len(sys.argv) > 1 and sys.argv[1] == "tc":
        run(1 << 14, [256] * 4, "tanh", reps=3)
        run(1 << 14, [128] * 4, "tanh", reps=3)
        sys.exit(0)
    run(1 << 16, [64] * 4, "tanh")
    run(1 << 16, [10, 20], "sigmoid")
    run(1 << 16, [16] * 4, "tanh")
    run(1 << 16, [10, 16], "sigmoid")
    run(6000, [20], "sigmoid", dim=1, inpDim=2, integNum=16, reps=50)
