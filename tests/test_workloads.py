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


def test_device_table_arguments_reproduce_the_host_table():
    """workloads.generate_on_device hands vn_generate_table_f64 the mesh centres and the periodic FE tables; the formulas
    that kernel evaluates (include/varnet_b200.h), restated here in NumPy float64, must give the float32 table that
    shard_feed builds on the host, bit for bit (the GPU test checks the kernel itself against the uploaded table)."""
    class Recorder:
        def generate_table(self, coord, tcoord, hVec, delta, N, dN, diff, vel, source, tf0, nb, integNum, integW, detJ):
            self.a = dict(coord=np.asarray(coord, float), tcoord=np.asarray(tcoord, float).ravel(), h=np.asarray(hVec, float).ravel(),
                          delta=np.asarray(delta, float), N=np.asarray(N, float).ravel(), dN=np.asarray(dN, float),
                          diff=diff, vel=np.asarray(vel, float), tf0=tf0, nb=nb, q=integNum, detJ=detJ, integW=integW)

    for integPnum in (2, 3):
        nx, ny, nti, n0, n1 = 5, 4, 3, 7, 41
        rec = Recorder()
        bic, meta = workloads.generate_on_device(rec, nx, ny, nti, n0, n1, integPnum=integPnum)
        a = rec.a
        feed, meta2 = workloads.shard_feed(nx, ny, nti, n0, n1, integPnum=integPnum, dtype=np.float32)
        assert meta == meta2 and a["nb"] == n1 - n0 and a["q"] == meta["integNum"]
        i = a["tf0"] + np.arange(a["nb"])
        s, j = i // nti, i % nti
        q = a["q"]
        dN = a["dN"].reshape(q, 3)
        X = np.stack([(a["coord"][s, d][:, None] + a["h"][d] * a["delta"][d][None, :]).reshape(-1) for d in range(2)] +
                     [(a["tcoord"][j][:, None] + a["h"][2] * a["delta"][2][None, :]).reshape(-1)], axis=1).astype(np.float32)
        G = np.tile(np.stack([a["diff"] * dN[:, k] + a["vel"][k] * a["N"] for k in range(2)], axis=1), [a["nb"], 1]).astype(np.float32)
        T = np.tile(dN[:, 2:3], [a["nb"], 1]).astype(np.float32)
        assert np.array_equal(X, feed["Input"]) and np.array_equal(G, feed["gcoef"]) and np.array_equal(T, feed["dNt"])
        assert a["detJ"] == feed["detJ"] and np.array_equal(bic["biInput"].astype(np.float32), feed["biInput"])
        assert bic["bDof"] == feed["bDof"] and bic["intShape"] == feed["intShape"]
