"""ncu target: one loss_grad on a small slice of the 4x64 tanh 2D+t workload (few waves per CTA)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from varnet_b200 import workloads
from varnet_b200._capi import Engine

lw = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "64,64,64,64".split(","))]
ntf = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 6
feed, meta = workloads.shard_feed(40, 40, 25, 0, ntf, dtype=np.float32)
rng = np.random.RandomState(0)
dims = [meta["inpDim"]] + lw + [1]
theta = np.concatenate([np.concatenate([rng.uniform(-1, 1, dims[i] * dims[i + 1]) * np.sqrt(6.0 / (dims[i] + dims[i + 1])), np.zeros(dims[i + 1])]) for i in range(len(dims) - 1)]).astype(np.float32)
eng = Engine(meta["dim"], meta["inpDim"], lw, "tanh", True, False, False, device=0)
eng.set_params(theta)
eng.upload_points(feed["Input"], feed["gcoef"], feed["source"], feed["N"], feed["dNt"], feed["intShape"], feed["integW"], feed["detJ"], False)
eng.upload_bic(feed["biInput"], feed["biLabel"], feed["bDof"], feed["biDimVal"])
eng.set_weights(feed["w"])
for _ in range(3):
    out = eng.loss_grad()
print("loss", out["loss"], eng.kernel_info())
eng.close()
