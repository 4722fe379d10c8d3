"""Phase cycle counters of the width-64 tensor-core tile kernel (one thread of CTA 0): where a tile's time goes.
usage: VARNET_B200_TC64_TIMING=1 python scripts/tc64_phases.py [ntf]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, ".")
os.environ.setdefault("VARNET_B200_TC64_TIMING", "1")
from varnet_b200 import workloads
from varnet_b200._capi import Engine, load_library

ntf = int(sys.argv[1]) if len(sys.argv) > 1 else 7104
lw = [64, 64, 64, 64]
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
t = (C.c_int64 * 16)()
rc = load_library().vn_debug_tc64_timing(t)
names = ["tile-start barrier (input prefetch landed)", "layer 0 (+ operands, stash)", "fwd epilogues", "fwd wait accumulator",
         "R_i / seeds / out-layer grads / first stash loads", "adj: operand store + signal + colsum + next zbar", "adj: wait weight-grad GEMM",
         "adj: drain + transposed stores (+ zbar of step 0)", "adj: wait layer GEMM", "adj: tail (last drain + waits)", "layer-0 gradients",
         "fold + loop"]
ntiles = (ntf * 64 // 128 + 147) // 148
tot = sum(t[i] for i in range(12))
print("rc", rc, "tiles of CTA 0:", ntiles, "cycles/tile: %.0f" % (tot / ntiles))
for i, n in enumerate(names):
    print("%-52s %9.0f cycles/tile  %5.1f %%" % (n, t[i] / ntiles, 100.0 * t[i] / tot))
eng.close()
