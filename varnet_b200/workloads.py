"""Named benchmark workloads (BASELINE.json `configs`) and a sharded feed generator.

`synthetic_2dt` is config 4/5: a rectangle discretised by the reference API into
nx*ny*ntime test functions ("space-time elements") x 4^3 Gauss points, constant kappa and
velocity, tanh MLP.  `shard_feed` builds the feed dict of one contiguous range of test functions
directly (what `VarNet.trainingPoints` + `trainData` + `ManageTrainData.trainDicts` would slice out
for that tower, VarNet.py:576-586,837 / VarNetUtility.py:830-854) without materialising the other
towers' rows, so that an 8-rank job does not build 8 copies of a 6.4e7-row table.
tests/test_workloads.py checks it against the host mirror bit for bit.
"""
import numpy as np

from .domain import PolygonDomain2D
from .fe import FE
from .hostutil import pair_rows
from .pde import ADPDE

SYNTH = dict(vertices=np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 1.0], [0.0, 1.0]]), diff=1.e-3, vel=[1., 0.],
             tInterval=[0, 1.0], bDiscNum=20)


def synthetic_pde():
    domain = PolygonDomain2D(SYNTH["vertices"])
    return ADPDE(domain, diff=SYNTH["diff"], vel=SYNTH["vel"], tInterval=SYNTH["tInterval"], IC=0.0)


def synthetic_2dt(VarNet, nx=100, ny=100, ntime=100, layerWidth=(64, 64, 64, 64), activation='tanh', **kw):
    return VarNet(synthetic_pde(), layerWidth=list(layerWidth), activationFun=activation, discNum=[nx, ny],
                  bDiscNum=SYNTH["bDiscNum"], tDiscNum=ntime, processors=kw.pop("processors", 'GPU:0'), **kw)


def tower_range(nt, puNum, tower, batchNum=1):
    """Contiguous test-function range of `tower` (VarNetUtility.py:821-838, batch 0)."""
    batchLen = int(np.ceil(nt / batchNum / puNum))
    n0 = min(tower * batchLen, nt)
    return n0, min(n0 + batchLen, nt)


def shard_feed(nx, ny, ntime, n0=None, n1=None, integPnum=2, dtype=np.float32, w=(1.0, 1.0, 1.0)):
    """Feed dict (reference keys) for test functions [n0, n1) of the synthetic 2D+t problem."""
    pde = synthetic_pde()
    domain = pde.domain
    mesh = domain.getMesh([nx, ny], SYNTH["bDiscNum"])
    t0, t1 = SYNTH["tInterval"]
    ht = (t1 - t0) / ntime
    t_coord = np.linspace(t0 + ht, t1, ntime).reshape(ntime, 1)
    hVec = np.vstack([np.reshape(mesh.he, [2, 1]), ht])
    per = FE(3, integPnum).periodic_tables(hVec)
    q = per["integNum"]
    nt = mesh.dof * ntime
    n0 = 0 if n0 is None else n0
    n1 = nt if n1 is None else n1
    idx = np.arange(n0, n1)
    s, j = idx // ntime, idx % ntime                                   # space index slow, time index fast
    delta = per["delta"]
    cols = [(mesh.coordinates[s, d][:, None] + mesh.he[d] * delta[d, :][None, :]).reshape(-1) for d in range(2)]
    cols.append((t_coord[j, 0][:, None] + ht * delta[2, :][None, :]).reshape(-1))
    Input = np.stack(cols, axis=1).astype(dtype)
    nb = n1 - n0
    N1 = per["N"].reshape(q, 1)
    dN = per["dN"].T                                                    # [q, 3]
    gco = SYNTH["diff"] * np.ones([q, 1]) * dN[:, 0:2] + np.asarray(SYNTH["vel"]) * np.ones([q, 2]) * N1
    # boundary rows of all 4 Dirichlet edges x time, then initial rows
    bInput = np.vstack([pair_rows(bc, t_coord) for bc in mesh.bCoordinates])
    iInput = np.concatenate([mesh.coordinates, np.zeros([mesh.dof, 1])], axis=1)
    biInput = np.vstack([bInput, iInput])
    feed = dict(Input=Input, gcoef=np.tile(gco, [nb, 1]).astype(dtype), source=np.zeros([nb * q, 1], dtype=dtype),
                N=np.tile(N1, [nb, 1]).astype(dtype), dNt=np.tile(dN[:, 2:3], [nb, 1]).astype(dtype),
                biInput=biInput.astype(dtype), biLabel=np.zeros([len(biInput), 1], dtype=dtype), bDof=len(bInput),
                intShape=[nb, q], integW=per["intWeight"], biDimVal=domain.measure, detJvec=False,
                detJ=per["detJ"], w=np.array(w, dtype=np.float64))
    meta = dict(nt=nt, integNum=q, dim=2, inpDim=3, timeDependent=True, lossOpt=dict(isSource=False, integWflag=integPnum != 2))
    return feed, meta


def generate_on_device(engine, nx, ny, ntime, n0=None, n1=None, integPnum=2):
    """The table of `shard_feed(nx, ny, ntime, n0, n1)` built on the device (vn_generate_table_f64): only the mesh
    centres and the periodic FE tables (a few KB) cross the bus; no nT-row host arrays are materialised.  Returns the
    boundary/initial part of the feed (small, still built on the host) and the same `meta` as `shard_feed`."""
    pde = synthetic_pde()
    domain = pde.domain
    mesh = domain.getMesh([nx, ny], SYNTH["bDiscNum"])
    t0, t1 = SYNTH["tInterval"]
    ht = (t1 - t0) / ntime
    t_coord = np.linspace(t0 + ht, t1, ntime).reshape(ntime, 1)
    hVec = np.vstack([np.reshape(mesh.he, [2, 1]), ht])
    per = FE(3, integPnum).periodic_tables(hVec)
    q = per["integNum"]
    nt = mesh.dof * ntime
    n0 = 0 if n0 is None else n0
    n1 = nt if n1 is None else n1
    engine.generate_table(mesh.coordinates, t_coord, hVec, per["delta"], per["N"], per["dN"].T, SYNTH["diff"], SYNTH["vel"], 0.0,
                          n0, n1 - n0, q, per["intWeight"], per["detJ"])
    bInput = np.vstack([pair_rows(bc, t_coord) for bc in mesh.bCoordinates])
    iInput = np.concatenate([mesh.coordinates, np.zeros([mesh.dof, 1])], axis=1)
    biInput = np.vstack([bInput, iInput])
    feed = dict(biInput=biInput, biLabel=np.zeros([len(biInput), 1]), bDof=len(bInput), biDimVal=domain.measure,
                intShape=[n1 - n0, q])
    meta = dict(nt=nt, integNum=q, dim=2, inpDim=3, timeDependent=True, lossOpt=dict(isSource=False, integWflag=integPnum != 2))
    return feed, meta


def mlp_macs(inpDim, layerWidth):
    dims = [inpDim] + list(layerWidth) + [1]
    return sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))


def algorithmic_flops_per_point(inpDim, dim, layerWidth):
    """F_alg = 6*(1+dim)*M flop per quadrature point for residual + gradient
    (forward value 2M, dim forward tangents 2M*dim, adjoint = 2x forward), M = sum in*out."""
    return 6 * (1 + dim) * mlp_macs(inpDim, layerWidth)
