"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
The fixtures travel to the GPU box, where /root/reference does not exist.

  fe_tables.npz         FE(dim, integPnum) tables + basisTot for dim 1..3, integPnum 2,3
                        (FiniteElement.py:55-434)
  feed_<config>.npz     the exact per-batch feed dict the reference's ManageTrainData builds
                        (VarNetUtility.py:771-868) for a scaled-down operator config, its index
                        tables, and the FP64 oracle's loss/gradient for a seeded weight vector.
"""
import os
import sys

import numpy as np

from . import configs
from . import graph_oracle as go
from .ref_loader import load_reference, reference_feed_dicts

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SCALES = dict(Operator_1Dt=(0.3, None), Operator_2Dt=(0.12, None), Operator_1DtMOR=(0.06, 4))


def fe_tables(ref):
    out = {}
    for dim in (1, 2, 3):
        for ip in (2, 3):
            fe = ref.FE(dim, ip)
            key = "d%d_p%d_" % (dim, ip)
            for name in ("basMultiInd", "IntegP", "basVal", "basDeriVal", "elemCoord", "delta"):
                out[key + name] = np.asarray(getattr(fe, name))
            if fe.IntegW is not None:
                out[key + "IntegW"] = fe.IntegW
            hVec = np.array([[0.09523809523809523], [0.024390243902439025], [0.02]])[:dim]
            integNum, nT, detJ, delta, iw, N, dN = fe.basisTot(5, hVec)
            out[key + "bt_scalars"] = np.array([integNum, nT, detJ])
            out[key + "bt_delta"], out[key + "bt_N"], out[key + "bt_dN"] = delta, N, dN
            if iw is not None:
                out[key + "bt_intWeight"] = iw
    return out


def feed_fixture(ref, name):
    scale, batchNum = SCALES[name]
    vn = configs.BUILDERS[name](ref, scale)
    tData, feeds, MORdiscArg = reference_feed_dicts(vn, batchNum=batchNum)
    tf = vn.tfData
    out = dict(scale=scale, batchNum=-1 if batchNum is None else batchNum, dim=tf.dim, inpDim=tf.inpDim,
               layerWidth=np.array(tf.layerWidth), timeDependent=int(tf.timeDependent),
               isSource=int(tf.lossOpt["isSource"]), integWflag=int(tf.lossOpt["integWflag"]),
               batchInd=tData.batchInd, integInd=tData.integInd, nbatch=len(feeds),
               hVec=vn.fixData.hVec, delta=vn.fixData.delta, nt=vn.fixData.nt,
               biDof=np.array(vn.fixData.biDof), uniform_input=vn.fixData.uniform_input)
    w = np.array([10.0, 10.0, 1.0])
    theta = go.glorot_init(tf.inpDim, tf.layerWidth, seed=2024)
    out["theta"] = theta
    out["w"] = w
    for b, towers in enumerate(feeds):
        fd = dict(towers[0]); fd["w"] = w
        for k, v in fd.items():
            if v is None:
                continue
            arr = np.asarray(v)
            if arr.dtype == object:
                continue
            out["b%d_%s" % (b, k)] = arr
        res = go.loss_and_grad(theta, fd, dim=tf.dim, inpDim=tf.inpDim, layerWidth=tf.layerWidth,
                               activation="sigmoid", timeDependent=tf.timeDependent, lossOpt=tf.lossOpt)
        out["b%d_oracle_scalars" % b] = np.array([res["loss"], res["BCloss"], res["ICloss"], res["varLoss"]])
        out["b%d_oracle_grad" % b] = res["grad"]
        out["b%d_oracle_lossVec" % b] = res["lossVec"]
        out["b%d_oracle_lossVec_tol" % b] = go.lossvec_tolerance(res)
        out["b%d_oracle_varLoss_tol" % b] = go.varloss_tolerance(res)
    return out


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "fe_tables.npz"), **fe_tables(ref))
    for name in SCALES:
        fx = feed_fixture(ref, name)
        path = os.path.join(OUT, "feed_%s.npz" % name)
        np.savez_compressed(path, **fx)
        print(name, "->", path, "%.1f KiB" % (os.path.getsize(path) / 1024), "batches", fx["nbatch"])


if __name__ == "__main__":
    sys.dont_write_bytecode = True
    main()
