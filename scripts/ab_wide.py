"""Same-box A/B of the tensor-core class (hidden widths 65..256) under different VARNET_B200_* switches, like scripts/ab_tc64.py:
  python scripts/ab_wide.py "name:VAR=1" "name2:" ...      each variant in its own process, ABAB.
A child prints ms per training step of depth-4 tanh networks of width 128 and 256 on a slice of the cfg-4 table (device
generated) and the worst relative gradient error against the FP64 oracle on a small slice."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    import numpy as np
    from varnet_b200 import workloads
    from varnet_b200._capi import Engine
    from oracle import graph_oracle as go
    from tests.util import rel_inf, layer_slices
    out = dict(name=name)
    for width in (128, 256):
        lw = [width] * 4
        theta = go.glorot_init(3, lw, seed=3)
        feed, meta = workloads.shard_feed(100, 100, 100, 400000, 400900, dtype=np.float64)
        kw = dict(dim=2, inpDim=3, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=meta["lossOpt"])
        ref = go.loss_and_grad(theta, feed, **kw)
        eng = Engine(2, 3, lw, "tanh", True, False, False, device=0)
        eng.set_params(theta)
        eng.upload_points(feed["Input"], feed["gcoef"], feed["source"], feed["N"], feed["dNt"], feed["intShape"], feed["integW"], feed["detJ"], False)
        eng.upload_bic(feed["biInput"], feed["biLabel"], feed["bDof"], feed["biDimVal"])
        eng.set_weights(feed["w"])
        o = eng.loss_grad()
        errs = [abs(float(o["loss"]) - ref["loss"]) / abs(ref["loss"])] + [float(rel_inf(o["grad"][sl], ref["grad"][sl])) for _, sl in layer_slices(3, lw)[:-1]]
        eng.close()
        ntf = int(os.environ.get("AB_NTF", "100000"))
        eng = Engine(2, 3, lw, "tanh", True, False, False, device=0)
        eng.set_params(theta)
        bfeed, meta = workloads.generate_on_device(eng, 100, 100, 100, 0, ntf)
        eng.upload_bic(bfeed["biInput"], bfeed["biLabel"], bfeed["bDof"], bfeed["biDimVal"])
        eng.set_weights(np.array([1.0, 1.0, 1.0]))
        eng.train_step(1e-3)
        eng.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            loss = eng.train_step(1e-3)
        eng.synchronize()
        dt = (time.perf_counter() - t0) / 3
        eng.close()
        out["w%d" % width] = dict(ms=round(dt * 1e3, 2), mpts=round(ntf * 64 / dt / 1e6, 2), worst_err=float("%.2e" % max(errs)), loss=float(loss))
    print(json.dumps(out), flush=True)


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2])
    variants = []
    for a in sys.argv[1:]:
        name, _, rest = a.partition(":")
        variants.append((name, dict(kv.split("=") for kv in rest.split(",") if kv)))
    for rep in range(2):
        for name, env in variants:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else ("FAILED %s: %s" % (name, (r.stdout + r.stderr)[-1500:])), flush=True)


if __name__ == "__main__":
    main()
