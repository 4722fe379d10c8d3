"""CPU: pins the oracle restatement (oracle/graph_oracle.py) — the reference ships no golden
vectors for loss/gradient, so it is checked against (i) an independent autograd
implementation mirroring the TF graph's tf.gradients structure and (ii) central finite differences."""
import numpy as np
import pytest

from oracle import graph_oracle as go
from oracle import torch_oracle as to
from tests.util import synth_feed, rel_inf

CASES = [
    (1, 2, [20], "sigmoid", True, False, False, False),
    (2, 3, [10, 20], "tanh", True, True, True, True),
    (2, 2, [8, 8, 8, 8], "tanh", False, True, False, False),
    (1, 3, [10, 20, 30], "sigmoid", True, False, True, False),
]


@pytest.mark.parametrize("case", CASES, ids=[str(c[2]) + c[3] for c in CASES])
def test_hand_adjoint_matches_autograd_double_backward(case):
    dim, inpDim, lw, act, td, src, iw, dvec = case
    rng = np.random.RandomState(1)
    feed = synth_feed(rng, dim, inpDim, 7, 16, 23, 15 if td else 23, td, src, iw, dvec)
    theta = go.glorot_init(inpDim, lw, seed=3) + 0.1 * rng.randn(go.param_count(inpDim, lw)).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation=act, timeDependent=td,
              lossOpt=dict(isSource=src, integWflag=iw))
    a = go.loss_and_grad(theta, feed, **kw)
    b = to.loss_and_grad(theta, feed, **kw)
    for k in ("loss", "BCloss", "ICloss", "varLoss"):
        assert abs(a[k] - b[k]) <= 1e-12 * max(1.0, abs(b[k]))
    assert rel_inf(a["grad"], b["grad"]) < 1e-12
    assert rel_inf(a["lossVec"], b["lossVec"]) < 1e-12


def test_gradient_matches_central_differences():
    rng = np.random.RandomState(2)
    dim, inpDim, lw = 1, 2, [3]
    feed = synth_feed(rng, dim, inpDim, 5, 4, 6, 4)
    # finite differences need the float32 feed rounding to be a no-op: use exactly representable weights
    theta = (np.round(rng.randn(go.param_count(inpDim, lw)) * 64) / 64).astype(np.float32)
    kw = dict(dim=dim, inpDim=inpDim, layerWidth=lw, activation="tanh", timeDependent=True,
              lossOpt=dict(isSource=False, integWflag=False))
    g = go.loss_and_grad(theta, feed, **kw)["grad"]
    h = 2.0 ** -10
    for i in range(theta.size):
        tp, tm = theta.copy(), theta.copy()
        tp[i] += h; tm[i] -= h
        fd = (go.loss_and_grad(tp, feed, need_grad=False, **kw)["loss"] -
              go.loss_and_grad(tm, feed, need_grad=False, **kw)["loss"]) / (2 * h)
        assert abs(fd - g[i]) <= 2e-4 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_tower_sum_equals_single_tower():
    """Splitting the test functions over towers and summing (TFModel.py:315-319,342-377) with the
    BC/IC weights divided by puNum (VarNetUtility.py:900-901) reproduces the single-tower result."""
    rng = np.random.RandomState(4)
    feed = synth_feed(rng, 1, 2, 12, 16, 20, 14)
    lw = [6]
    theta = go.glorot_init(2, lw, seed=1)
    kw = dict(dim=1, inpDim=2, layerWidth=lw, activation="sigmoid", timeDependent=True,
              lossOpt=dict(isSource=False, integWflag=False))
    one = go.loss_and_grad(theta, feed, **kw)
    towers = []
    for lo, hi in ((0, 7), (7, 12)):
        f = dict(feed)
        sl = slice(lo * 16, hi * 16)
        for k in ("Input", "gcoef", "source", "N", "dNt"):
            f[k] = feed[k][sl]
        f["intShape"] = [hi - lo, 16]
        f["w"] = feed["w"] * np.array([0.5, 0.5, 1.0])
        towers.append(f)
    two = go.towers_loss_and_grad(theta, towers, **kw)
    assert abs(two["loss"] - one["loss"]) < 1e-12 * abs(one["loss"])
    assert rel_inf(two["grad"], one["grad"]) < 1e-12
    assert np.allclose(two["lossVec"], one["lossVec"], rtol=1e-14)


def test_strong_residual_against_autograd():
    import torch
    rng = np.random.RandomState(6)
    dim, inpDim, lw = 2, 3, [7, 5]
    theta = go.glorot_init(inpDim, lw, seed=5)
    X = rng.uniform(-1, 1, (9, inpDim)).astype(np.float32)
    diff = rng.rand(9, 1); vel = rng.randn(9, 2); ddx = rng.randn(9, 2); src = rng.randn(9, 1)
    u, res = go.strong_residual(theta, X, diff, vel, ddx, src, dim, inpDim, lw, "tanh", True)
    th = torch.tensor(theta.astype(np.float64))
    Ws, bs = to._split(th, inpDim, lw)
    Xt = torch.tensor(X.astype(np.float64), requires_grad=True)
    ut = to._model(Xt, Ws, bs, go.ACT_TANH)
    g = torch.autograd.grad(ut.sum(), Xt, create_graph=True)[0]
    lap = sum(torch.autograd.grad(g[:, d].sum(), Xt, retain_graph=True)[0][:, d] for d in range(dim))
    f32 = lambda a: torch.tensor(np.asarray(a, dtype=np.float32).astype(np.float64))
    ref = -g[:, dim] + f32(diff)[:, 0] * lap - ((f32(vel) - f32(ddx)) * g[:, :dim]).sum(-1) + f32(src)[:, 0]
    assert rel_inf(res, ref.detach().numpy()) < 1e-12
    assert rel_inf(u, ut.detach().numpy()[:, 0]) < 1e-13


def test_tf_adam_formula():
    th = np.array([1.0, -2.0]); g = np.array([0.5, -0.25])
    m = np.zeros(2); v = np.zeros(2)
    th1, m1, v1 = go.adam_step(th, g, m, v, 1)
    # first TF-Adam step: lr*sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + eps)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = th - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8)
    assert np.allclose(th1, exp, rtol=0, atol=1e-15)
