"""Same-box A/B of tc64_var_kernel variants: every variant is a set of VARNET_B200_* environment variables, run in its own
process (the switches are read once per process) on the same GPU, back to back, twice (ABAB) so that drift shows.

  python scripts/ab_tc64.py "name:VAR=1,VAR2=3" "name2:" ...          (driver)
  python scripts/ab_tc64.py --child name                              (one measurement)

A child prints: kernel ms of a 1/4 cfg-4 launch (250 000 test functions x 64 Gauss points, 4x64 tanh, table generated on the
device), M quad-pts/s, and the worst relative error of loss / gradient tensors against the FP64 oracle on a 3 000-test-function
slice of the same table (uploaded), so that a faster variant that is less accurate is seen at once."""
import os
import subprocess
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name):
    import numpy as np
    from varnet_b200 import workloads
    from varnet_b200._capi import Engine
    from oracle import graph_oracle as go
    from tests.util import rel_inf, layer_slices
    lw = [64] * 4
    theta = go.glorot_init(3, lw, seed=3)
    # accuracy on an oracle-sized slice (interior range, ragged tile count)
    feed, meta = workloads.shard_feed(100, 100, 100, 400000, 403001, dtype=np.float64)
    kw = dict(dim=2, inpDim=3, layerWidth=lw, activation="tanh", timeDependent=True, lossOpt=meta["lossOpt"])
    ref = go.loss_and_grad(theta, feed, **kw)
    eng = Engine(2, 3, lw, "tanh", True, False, False, device=0)
    eng.set_params(theta)
    eng.upload_points(feed["Input"], feed["gcoef"], feed["source"], feed["N"], feed["dNt"], feed["intShape"], feed["integW"], feed["detJ"], False)
    eng.upload_bic(feed["biInput"], feed["biLabel"], feed["bDof"], feed["biDimVal"])
    eng.set_weights(feed["w"])
    out = eng.loss_grad()
    errs = {"loss": abs(float(out["loss"]) - ref["loss"]) / abs(ref["loss"])}
    for nm, sl in layer_slices(3, lw)[:-1]:
        errs[nm] = float(rel_inf(out["grad"][sl], ref["grad"][sl]))
    worst = max(errs.values())
    info = eng.kernel_info()
    eng.close()
    # speed
    ntf = int(os.environ.get("AB_NTF", "250000"))
    eng = Engine(2, 3, lw, "tanh", True, False, False, device=0)
    eng.set_params(theta)
    if os.environ.get("AB_UPLOAD"):             # streamed (uploaded) table instead of the in-kernel generated one
        bfeed, meta = workloads.shard_feed(100, 100, 100, 0, ntf, dtype=np.float32)
        eng.upload_points(bfeed["Input"], bfeed["gcoef"], bfeed["source"], bfeed["N"], bfeed["dNt"], bfeed["intShape"], bfeed["integW"], bfeed["detJ"], False)
    else:
        bfeed, meta = workloads.generate_on_device(eng, 100, 100, 100, 0, ntf)
    eng.upload_bic(bfeed["biInput"], bfeed["biLabel"], bfeed["bDof"], bfeed["biDimVal"])
    eng.set_weights(np.array([1.0, 1.0, 1.0]))
    g0 = eng.loss_grad()["grad"].astype(np.float64)          # gradient of the large launch (window folds exercised)
    np.save("/tmp/ab_grad_%s.npy" % name, g0)
    gdiff = None
    if os.environ.get("AB_REF") and os.path.exists("/tmp/ab_grad_%s.npy" % os.environ["AB_REF"]):
        gr = np.load("/tmp/ab_grad_%s.npy" % os.environ["AB_REF"])
        gdiff = max(float(rel_inf(g0[sl], gr[sl])) for _, sl in layer_slices(3, lw)[:-1])
    for _ in range(3):
        eng.train_step(1e-3)
    eng.profile_enable(True); eng.profile_read()
    for _ in range(6):
        loss = eng.train_step(1e-3)
    pr = eng.profile_read()
    k = pr["var_adj"][0] / pr["var_adj"][1]
    eng.close()
    print(json.dumps(dict(name=name, kernel_ms=round(k, 3), mpts=round(ntf * 64 / k / 1e3, 1), worst_err=worst, grad_vs_first_variant=gdiff, loss=float(loss),
                          errs={a: float("%.2e" % b) for a, b in errs.items()}, info=info.split("|")[0][:80])), flush=True)


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2])
    variants = []
    for a in sys.argv[1:]:
        name, _, rest = a.partition(":")
        variants.append((name, dict(kv.split("=") for kv in rest.split(",") if kv)))
    for rep in range(2):
        for name, env in variants:
            e = dict(os.environ, AB_REF=variants[0][0], **env)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name], env=e, capture_output=True, text=True, timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(line[-1] if line else ("FAILED %s: %s" % (name, (r.stdout + r.stderr)[-1500:])), flush=True)


if __name__ == "__main__":
    main()
