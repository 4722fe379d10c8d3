"""End-to-end pin on an answer the REFERENCE itself holds.

The reference ships no loss/gradient vectors, but its operator script for config 1 carries the analytic solution of the
problem it trains on — a Fourier series (`/root/reference/Operator_1Dt.py:78-110`, `cExact`) — and measures the trained
network against it with `uf.l2Err(cEx, cApp)` (`Operator_1Dt.py:178-187`, `UtilityFunc.py` l2Err = ||cEx-cApp|| / ||cEx||).
This test does the same through the drop-in API: `VarNet.train` (uniform sampling, the reference's loss weights
[10, 10, 1]) on the FULL-size config 1 (20 x 300 test functions, 96 000 quadrature points, 1 x 20 sigmoid network) on the
GPU, then `l2Err` of `evaluate()` on the reference's own evaluation grid (`FIXData.setInputData`, uniform_input).
A loss or gradient that deviated from the reference's graph could not drive the network to the analytic solution."""
import tempfile

import numpy as np
import pytest
from numpy import pi, sin, cos, exp



U, D, T = 1.0, 0.1 / pi, 2.0          # Operator_1Dt.py:71-73


def IC(x):                             # Operator_1Dt.py:76-77
    return -sin(pi * x)


def cExact(x, t, trunc=800):
    """Restated from Operator_1Dt.py:79-110 (same truncation, same formula)."""
    ind0 = t == 0
    cInit = IC(x[ind0])
    p = np.arange(0, trunc + 1.0).reshape(1, trunc + 1)
    c0 = 16 * pi ** 2 * D ** 3 * U * exp(U / D / 2 * (x - U * t / 2))
    c1_n = (-1) ** p * 2 * p * sin(p * pi * x) * exp(-D * p ** 2 * pi ** 2 * t)
    c1_d = U ** 4 + 8 * (U * pi * D) ** 2 * (p ** 2 + 1) + 16 * (pi * D) ** 4 * (p ** 2 - 1) ** 2
    c1 = np.sinh(U / D / 2) * np.sum(c1_n / c1_d, axis=-1, keepdims=True)
    c2_n = (-1) ** p * (2 * p + 1) * cos((p + 0.5) * pi * x) * exp(-D * (2 * p + 1) ** 2 * pi ** 2 * t / 4)
    c2_d = U ** 4 + (U * pi * D) ** 2 * (8 * p ** 2 + 8 * p + 10) + (pi * D) ** 4 * (4 * p ** 2 + 4 * p - 3) ** 2
    c2 = np.cosh(U / D / 2) * np.sum(c2_n / c2_d, axis=-1, keepdims=True)
    c = c0 * (c1 + c2)
    c[ind0] = cInit
    return c


EPOCHS = 60000
L2ERR_BOUND = 0.10            # achieved on a B200 with seed 1: see the printed value (set from the measured trajectory)


def build():
    import varnet_b200
    domain = varnet_b200.Domain1D()
    pde = varnet_b200.ADPDE(domain, diff=D, vel=U, timeDependent=True, tInterval=[0, T], IC=IC, cEx=cExact)
    return varnet_b200.VarNet(pde, layerWidth=[20], discNum=20, bDiscNum=None, tDiscNum=300, processors='GPU:0', seed=1)


def test_series_solution_is_the_pde_solution():
    """Sanity of the restated series: zero Dirichlet values at x = +-1, initial condition at t = 0, and the PDE
    c_t + u c_x = D c_xx by central differences at interior points."""
    t = np.full((5, 1), 0.7)
    assert np.abs(cExact(np.full((5, 1), 1.0), t)).max() < 1e-6 and np.abs(cExact(np.full((5, 1), -1.0), t)).max() < 1e-6
    x = np.linspace(-0.9, 0.9, 7).reshape(-1, 1)
    assert np.allclose(cExact(x, np.zeros_like(x)), IC(x))
    tt = np.full_like(x, 0.9)
    h = 1e-3
    c_t = (cExact(x, tt + h) - cExact(x, tt - h)) / (2 * h)
    c_x = (cExact(x + h, tt) - cExact(x - h, tt)) / (2 * h)
    c_xx = (cExact(x + h, tt) - 2 * cExact(x, tt) + cExact(x - h, tt)) / h ** 2
    assert np.abs(c_t + U * c_x - D * c_xx).max() < 1e-4 * np.abs(c_t).max() + 1e-5


@pytest.mark.gpu
def test_operator_1dt_trains_to_the_analytic_solution():
    from varnet_b200.hostutil import l2_err
    np.random.seed(0)
    vn = build()
    cEx = vn.fixData.cEx
    err0 = l2_err(cEx, vn.evaluate())
    with tempfile.TemporaryDirectory() as d:
        res = vn.train(d, weight=[10., 10., 1.], smpScheme='uniform', epochNum=EPOCHS, saveFreq=5000, verbose=False, tol=1e-1)
        traj = [float(e) for e in res.error if e is not None]
        vn.loadModel()
        err = l2_err(cEx, vn.evaluate())
    print("[achieved] Operator_1Dt end-to-end: l2Err %.4f -> %.4f after %d epochs; loss %.4g -> %.4g; error every 5000 epochs: %s"
          % (err0, err, len(res.loss), res.loss[0], res.loss[-1], ["%.3f" % e for e in traj]))
    assert err < L2ERR_BOUND, (err0, err, traj)
    assert err < 0.25 * err0
    vn.tfData.sess.close()
