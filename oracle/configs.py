"""TEST INFRASTRUCTURE — the reference's operator configurations as plain builders.

`build(name, api, scale)` constructs the PDE + VarNet of a BASELINE.json config with a given
API namespace: `api` is either the reference modules (oracle.ref_loader.load_reference()) or
`varnet_b200`, so the very same script drives both sides of a table-parity test.
Parameters: Operator_1Dt.py:69-160, Operator_2Dt.py:80-157, Operator_1DtMOR.py:68-204.
`scale` < 1 shrinks the discretisation for fast tests (the structure is unchanged).
"""
import numpy as np

pi = np.pi


def _n(v, scale, lo=2):
    return max(lo, int(round(v * scale)))


def operator_1dt(api, scale=1.0, bDiscNum=None, **kw):
    domain = api.Domain1D()
    pde = api.ADPDE(domain, diff=0.1 / pi, vel=1.0, timeDependent=True, tInterval=[0, 2.0],
                    IC=lambda x: -np.sin(pi * x))
    return api.VarNet(pde, layerWidth=[20], discNum=_n(20, scale), bDiscNum=bDiscNum, tDiscNum=_n(300, scale),
                      processors='GPU:0', **kw)


def operator_2dt(api, scale=1.0, **kw):
    vertices = np.array([[0.0, -0.5], [0.0, -0.2], [0.0, 0.2], [0.0, 0.5], [2.0, 0.5], [2.0, -0.5]])
    domain = api.PolygonDomain2D(vertices)
    BC = [[], [0.0, 1.0, 1.0], [], [], [], []]
    pde = api.ADPDE(domain, diff=1.e-3, vel=[1., 0.], tInterval=[0, 1.5], BCs=BC, IC=0.0)
    return api.VarNet(pde, layerWidth=[10, 20], discNum=[_n(80, scale), _n(40, scale)], bDiscNum=_n(40, scale, 4),
                      tDiscNum=_n(75, scale), processors='GPU:0', **kw)


def diffFun(x, t=0, D=0.1 / pi):
    return D * np.ones([np.shape(x)[0], 1])


def discDiff(discNum=6):
    return np.array([0.003 * (11 ** (n / (discNum - 1))) for n in range(discNum)])[np.newaxis].T


def operator_1dtmor(api, scale=1.0, **kw):
    mor = api.MOR(diffFun, ['D'], [[0.003, 0.033]])
    domain = api.Domain1D()
    pde = api.ADPDE(domain, diff=diffFun, vel=1.0, timeDependent=True, tInterval=[0, 2.0],
                    IC=lambda x: -np.sin(pi * x), MORvar=mor)
    return api.VarNet(pde, layerWidth=[10, 20, 30], discNum=_n(150, scale), bDiscNum=75, tDiscNum=_n(800, scale),
                      MORdiscScheme=discDiff, processors='GPU:0', **kw)


def synthetic_2dt(api, nx=100, ny=100, ntime=100, layerWidth=(64, 64, 64, 64), activation='tanh', **kw):
    """Config 4/5: rectangle, constant kappa / vel, nt = nx*ny*ntime test functions x 4^3 Gauss points."""
    domain = api.PolygonDomain2D(np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 1.0], [0.0, 1.0]]))
    pde = api.ADPDE(domain, diff=1.e-3, vel=[1., 0.], tInterval=[0, 1.0], IC=0.0)
    return api.VarNet(pde, layerWidth=list(layerWidth), activationFun=activation, discNum=[nx, ny], bDiscNum=20,
                      tDiscNum=ntime, processors='GPU:0', **kw)


BUILDERS = dict(Operator_1Dt=operator_1dt, Operator_2Dt=operator_2dt, Operator_1DtMOR=operator_1dtmor)
TRAIN_KW = dict(Operator_1Dt=dict(), Operator_2Dt=dict(), Operator_1DtMOR=dict(batchNum=20))
