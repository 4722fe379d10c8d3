"""CPU: the sharded synthetic-workload generator equals the slice the host mirror (and hence the
reference's trainDicts, VarNetUtility.py:830-854) would feed to that tower."""
import types

import numpy as np

import varnet_b200
from varnet_b200 import workloads
from oracle.ref_loader import RecorderTFNN


def test_shard_feed_equals_mirror_slices(monkeypatch):
    import varnet_b200.trainer as tr
    monkeypatch.setattr(tr, "TFNN", lambda *a, seed=None, **k: RecorderTFNN(*a, **k))
    nx, ny, nti = 6, 5, 4
    vn = workloads.synthetic_2dt(tr.VarNet, nx, ny, nti, layerWidth=(8, 8))
    fd = vn.fixData
    fd.setFEdata()
    Input, _, biInput, _ = vn.trainingPoints()
    tData = varnet_b200.ManageTrainData(Input, biInput, None, None, False, 1)
    tData = vn.trainData(0, None, tData)
    full = {k.split("/")[1]: v for k, v in tData.optimFeedicts[0].items()}
    nt = fd.nt
    assert nt == nx * ny * nti
    for world in (1, 3):
        for rank in range(world):
            n0, n1 = workloads.tower_range(nt, world, rank)
            feed, meta = workloads.shard_feed(nx, ny, nti, n0, n1, dtype=np.float64)
            q = meta["integNum"]
            sl = slice(n0 * q, n1 * q)
            for k in ("Input", "gcoef", "N", "dNt", "source"):
                assert np.array_equal(feed[k], full[k][sl]), (world, rank, k)
            for k in ("biInput", "biLabel"):
                assert np.array_equal(feed[k], full[k]), k
            assert feed["bDof"] == full["bDof"] and feed["intShape"] == [n1 - n0, q]
            assert feed["detJ"] == full["detJ"] and feed["biDimVal"] == full["biDimVal"]
            assert meta["lossOpt"] == vn.tfData.lossOpt


def test_tower_ranges_cover_everything_like_the_reference():
    for nt, pu in ((1000000, 8), (6000, 1), (10, 3), (7, 8)):
        covered = []
        for r in range(pu):
            n0, n1 = workloads.tower_range(nt, pu, r)
            covered.extend(range(n0, n1))
        assert covered == list(range(nt))


def test_algorithmic_flops_match_survey_table():
    assert workloads.algorithmic_flops_per_point(2, 1, [20]) == 720
    assert workloads.algorithmic_flops_per_point(3, 2, [10, 20]) == 4500
    assert workloads.algorithmic_flops_per_point(3, 1, [10, 20, 30]) == 10320
    assert workloads.algorithmic_flops_per_point(3, 2, [64] * 4) == 225792
