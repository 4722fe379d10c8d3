"""Per-tile time of the cfg-4 kernel against the number of tiles per launch (one GPU): is the N-rank loss a size effect?"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from varnet_b200 import workloads
from varnet_b200._capi import Engine
from oracle import graph_oracle as go

lw = [64] * 4
theta = go.glorot_init(3, lw, seed=3)
for ntf in (31250, 62500, 125000, 250000, 500000):
    feed, meta = workloads.shard_feed(100, 100, 100, 0, ntf, dtype=np.float32)
    eng = Engine(2, 3, lw, "tanh", True, False, False, device=0)
    eng.set_params(theta)
    eng.upload_points(feed["Input"], feed["gcoef"], feed["source"], feed["N"], feed["dNt"], feed["intShape"], feed["integW"], feed["detJ"], False)
    eng.upload_bic(feed["biInput"], feed["biLabel"], feed["bDof"], feed["biDimVal"])
    eng.set_weights(feed["w"])
    for _ in range(3):
        eng.train_step(1e-3)
    eng.profile_enable(True); eng.profile_read()
    for _ in range(5):
        eng.train_step(1e-3)
    pr = eng.profile_read()
    k = pr["var_adj"][0] / pr["var_adj"][1]
    tiles = ntf * 64 // 128
    print("ntf %7d  tiles/CTA %7.1f  kernel %8.3f ms  per tile-round %.2f us  M pts/s %.1f" % (ntf, tiles / 148, k, k * 1e3 / np.ceil(tiles / 148), ntf * 64 / k / 1e3))
    eng.close()
