"""TEST INFRASTRUCTURE — the parity oracle, not product code.

CPU restatement (NumPy, float64 by default) of the reference's TensorFlow-1.10
training graph for the weak-form advection-diffusion loss:

  * model            TFNN.defModel            /root/reference/TFModel.py:195-249
  * input gradients  NNModel.modelGrad        /root/reference/TFModel.py:515-564
  * loss             NNModel.LossFun          /root/reference/TFModel.py:567-691
  * weight gradient  NNModel.computeGrad      /root/reference/TFModel.py:695-714
  * tower sum        TFNN.sum_grads/optimSetup /root/reference/TFModel.py:293-377
  * strong residual  NNModel.Residual         /root/reference/TFModel.py:718-772
  * optimizers       tf.train.AdamOptimizer / RMSPropOptimizer (TF 1.10, un-vendored
                     third-party dependency; published update rules restated below)

PARITY PINNING: the arithmetic itself lives in TensorFlow 1.10 + tf.keras
(README.md:20-25 of the reference), which is absent from /root/reference and cannot
be installed here, and the reference ships no tests/golden vectors for loss or
gradients => **"parity unpinned" at the TF boundary**.  What *is*
pinned: (i) every input table, bit-exactly, by running the reference's own NumPy
code (oracle/ref_loader.py -> tests/golden/*.npz); (ii) this restatement against an
independent automatic-differentiation implementation that mirrors the graph's
tf.gradients structure (oracle/torch_oracle.py, float64 double-backward) and against
central finite differences (tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.

Flat parameter layout (Keras trainable-variable order): for each Dense layer
kernel[in,out] row-major then bias[out]; last layer is Dense(1) 'output'.
"""
import numpy as np

ACT_SIGMOID = 0
ACT_TANH = 1
_ACT_IDS = {"sigmoid": ACT_SIGMOID, "tanh": ACT_TANH}


def act_id(name):
    if isinstance(name, (list, tuple)):
        ids = {act_id(n) for n in name}
        if len(ids) != 1:
            raise ValueError("mixed activation lists are not supported by the oracle")
        return ids.pop()
    if isinstance(name, (int, np.integer)):
        return int(name)
    return _ACT_IDS[name.lower()]


def layer_sizes(inpDim, layerWidth):
    """[(in,out)] for hidden layers + the Dense(1) output (TFModel.py:210-242)."""
    dims = [inpDim] + list(layerWidth) + [1]
    return [(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]


def param_count(inpDim, layerWidth):
    return sum(i * o + o for i, o in layer_sizes(inpDim, layerWidth))


def unpack(theta, inpDim, layerWidth):
    Ws, bs, off = [], [], 0
    for i, o in layer_sizes(inpDim, layerWidth):
        Ws.append(theta[off:off + i * o].reshape(i, o)); off += i * o
        bs.append(theta[off:off + o]); off += o
    assert off == theta.size
    return Ws, bs


def pack(Ws, bs):
    return np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(Ws, bs)])


def glorot_init(inpDim, layerWidth, seed=0, dtype=np.float32):
    """glorot_uniform kernels, zero biases (TFModel.py:210-242).  The reference never
    seeds, so any fixed seed is a legitimate initialisation."""
    rng = np.random.RandomState(seed)
    Ws, bs = [], []
    for i, o in layer_sizes(inpDim, layerWidth):
        lim = np.sqrt(6.0 / (i + o))
        Ws.append(rng.uniform(-lim, lim, size=(i, o)))
        bs.append(np.zeros(o))
    return pack(Ws, bs).astype(dtype)


def _act(z, act):
    if act == ACT_SIGMOID:
        return 1.0 / (1.0 + np.exp(-z))
    return np.tanh(z)


def _d1(a, act):            # act'(z) expressed in a = act(z)
    return a * (1.0 - a) if act == ACT_SIGMOID else 1.0 - a * a


def _d2_over_d1(a, act):    # act''/act'
    return (1.0 - 2.0 * a) if act == ACT_SIGMOID else -2.0 * a


def _d3_terms(a, act):
    """act''' expressed in a (needed only by the strong-form residual's Laplacian)."""
    if act == ACT_SIGMOID:
        s1 = a * (1 - a)
        return s1 * (1 - 6 * a + 6 * a * a)
    t1 = 1 - a * a
    return t1 * (6 * a * a - 2)


def mlp_forward(theta, X, inpDim, layerWidth, act, ndir):
    """Value and forward-mode input tangents for the first ``ndir`` input columns.
    Equals model(Input) and tf.gradients(model(Input), Input)[0][:, :ndir]
    (TFModel.py:536-541).  Returns u[P], du[ndir,P], caches (A, dA)."""
    Ws, bs = unpack(theta, inpDim, layerWidth)
    L = len(layerWidth)
    a = X
    A, dA = [], []
    da = None
    for l in range(L):
        z = a @ Ws[l] + bs[l]
        if l == 0:
            dz = np.broadcast_to(Ws[0][:ndir, None, :], (ndir,) + z.shape)
        else:
            dz = (da.reshape(-1, da.shape[-1]) @ Ws[l]).reshape(ndir, z.shape[0], -1)    # one BLAS call for all tangents
        a = _act(z, act)
        da = _d1(a, act)[None] * dz
        A.append(a); dA.append(da)
    u = a @ Ws[L][:, 0] + bs[L][0]
    du = da @ Ws[L][:, 0]
    return u, du, (A, dA)


def mlp_value(theta, X, inpDim, layerWidth, act):
    Ws, bs = unpack(theta, inpDim, layerWidth)
    a = X
    for l in range(len(layerWidth)):
        a = _act(a @ Ws[l] + bs[l], act)
    return a @ Ws[-1][:, 0] + bs[-1][0]


def _cast_feed(feed, dtype):
    """The TF placeholders are float32: every fed array is rounded to float32 before
    any arithmetic (TFModel.py:531,602-620).  ``dtype`` is the arithmetic dtype after
    that rounding (float64 for the parity oracle, float32 for the timed CPU port)."""
    def c(v):
        return np.asarray(v, dtype=np.float32).astype(dtype)
    out = dict(feed)
    for k in ("Input", "biInput", "biLabel", "gcoef", "source", "N", "dNt", "detJ", "integW",
              "biDimVal", "w", "diff", "vel", "diff_dx"):
        v = feed.get(k)
        if v is None:
            continue
        arr = np.asarray(v, dtype=object) if isinstance(v, list) else np.asarray(v)
        if arr.dtype == object or arr.size == 0 or (arr.dtype.kind not in "fiub"):
            out[k] = None          # e.g. dNt=[[None]] for steady problems (VarNetUtility.py:449)
        else:
            out[k] = c(v)
    return out


def loss_and_grad(theta, feed, dim, inpDim, layerWidth, activation, timeDependent, lossOpt,
                  need_grad=True, dtype=np.float64):
    """One tower of the reference graph (TFModel.py:567-714) on one feed dict.

    feed keys = the reference's per-tower feed keys (VarNetUtility.py:840-854):
    Input, biInput, biLabel, gcoef, source, N, dNt, bDof, intShape, integW, biDimVal,
    detJvec, detJ, w.
    Returns dict(loss, BCloss, ICloss, varLoss, lossVec[nb], R[nb], grad[flat] | None).
    """
    act = act_id(activation)
    f = _cast_feed(feed, dtype)
    theta = np.asarray(theta, dtype=np.float32).astype(dtype)
    Ws, bs = unpack(theta, inpDim, layerWidth)
    L = len(layerWidth)
    X = f["Input"].reshape(-1, inpDim)
    nb, integNum = [int(v) for v in feed["intShape"]]
    P = nb * integNum
    assert X.shape[0] == P
    gcoef = f["gcoef"].reshape(P, dim)
    w = f["w"].reshape(3)
    detJvec = bool(feed.get("detJvec", False))
    detJ = f["detJ"].reshape(-1)                       # scalar or [nb]
    biDimVal = float(f["biDimVal"])

    # ---- variational term (TFModel.py:653-664)
    u, du, (A, dA) = mlp_forward(theta, X, inpDim, layerWidth, act, dim)
    I = np.einsum("kp,pk->p", du, gcoef)
    Iabs = np.einsum("kp,pk->p", np.abs(du), np.abs(gcoef))      # magnitude of the summed terms (conditioning)
    if timeDependent:
        dNt = f["dNt"].reshape(P)
        I = I - u * dNt
        Iabs = Iabs + np.abs(u * dNt)
    if lossOpt["isSource"]:
        I = I - f["source"].reshape(P) * f["N"].reshape(P)
        Iabs = Iabs + np.abs(f["source"].reshape(P) * f["N"].reshape(P))
    I2 = I.reshape(nb, integNum)
    if lossOpt["integWflag"]:
        wq = f["integW"].reshape(1, integNum)
        I2 = wq * I2
    else:
        wq = np.ones((1, integNum), dtype=dtype)
    R = I2.sum(axis=1)
    Rabs = (np.abs(wq) * Iabs.reshape(nb, integNum)).sum(axis=1)   # sum_q |w_q| * (|terms| of I_q)
    R2 = R * R
    if detJvec:
        varLoss = np.sum(detJ * R2)
    else:
        varLoss = detJ[0] * np.sum(R2)
    lossVec = detJ * R2                                 # TFModel.py:668

    # ---- boundary / initial term (TFModel.py:643-650)
    bX = f["biInput"].reshape(-1, inpDim)
    bL = f["biLabel"].reshape(-1)
    bDof = int(feed["bDof"])
    nbi = bX.shape[0]
    Ws_b, bs_b = Ws, bs
    ab = bX
    Ab = []
    for l in range(L):
        ab = _act(ab @ Ws[l] + bs[l], act)
        Ab.append(ab)
    ub = ab @ Ws[L][:, 0] + bs[L][0]
    r = ub - bL
    biCs = biDimVal * r * r
    with np.errstate(invalid="ignore", divide="ignore"):
        bCs = biCs[:bDof].mean() if bDof > 0 else np.float64(np.nan)
        if timeDependent:
            iCs = biCs[bDof:].mean() if nbi > bDof else np.float64(np.nan)   # tf.reduce_mean of empty -> nan
        else:
            iCs = np.float64(0.0)
    loss = w[0] * bCs + w[1] * iCs + w[2] * varLoss
    out = dict(loss=loss, BCloss=bCs, ICloss=iCs, varLoss=varLoss, lossVec=lossVec, R=R, Rabs=Rabs,
               detJ=np.broadcast_to(detJ, (nb,)) if detJvec else np.full(nb, detJ[0]), grad=None)
    if not need_grad:
        return out

    # ---- adjoint: d loss / d theta
    gW = [np.zeros_like(W) for W in Ws]
    gb = [np.zeros_like(b) for b in bs]
    dJ = detJ if detJvec else detJ[0]
    lam = (2.0 * w[2] * dJ * R)[:, None] * wq           # [nb, integNum]
    lam = lam.reshape(P)
    ubar = -lam * dNt if timeDependent else np.zeros(P, dtype=dtype)
    ukbar = lam[None, :] * gcoef.T                       # [dim,P]
    aL, daL = A[L - 1], dA[L - 1]
    gW[L][:, 0] += aL.T @ ubar + daL.reshape(-1, daL.shape[-1]).T @ ukbar.reshape(-1)
    gb[L][0] += ubar.sum()
    abar = ubar[:, None] * Ws[L][:, 0][None, :]
    dabar = ukbar[:, :, None] * Ws[L][:, 0][None, None, :]
    for l in range(L - 1, -1, -1):
        a, da = A[l], dA[l]
        d1 = _d1(a, act)
        dzbar = dabar * d1[None]
        zbar = abar * d1 + _d2_over_d1(a, act) * (dabar * da).sum(axis=0)
        if l > 0:
            ap, dap = A[l - 1], dA[l - 1]
            gW[l] += ap.T @ zbar + dap.reshape(-1, dap.shape[-1]).T @ dzbar.reshape(-1, dzbar.shape[-1])
            abar = zbar @ Ws[l].T
            dabar = (dzbar.reshape(-1, dzbar.shape[-1]) @ Ws[l].T).reshape(dzbar.shape[0], dzbar.shape[1], -1)
        else:
            gW[0] += X.T @ zbar
            gW[0][:dim, :] += dzbar.sum(axis=1)
        gb[l] += zbar.sum(axis=0)

    # BC/IC: plain back-prop with per-row seeds
    seed = np.zeros(nbi, dtype=dtype)
    if bDof > 0:
        seed[:bDof] = w[0] / bDof
    if timeDependent and nbi > bDof:
        seed[bDof:] = w[1] / (nbi - bDof)
    ubar_b = 2.0 * biDimVal * r * seed
    gW[L][:, 0] += Ab[L - 1].T @ ubar_b
    gb[L][0] += ubar_b.sum()
    abar = ubar_b[:, None] * Ws[L][:, 0][None, :]
    for l in range(L - 1, -1, -1):
        zbar = abar * _d1(Ab[l], act)
        prev = Ab[l - 1] if l > 0 else bX
        gW[l] += prev.T @ zbar
        gb[l] += zbar.sum(axis=0)
        if l > 0:
            abar = zbar @ Ws[l].T
    out["grad"] = pack(gW, gb)
    return out


def towers_loss_and_grad(theta, tower_feeds, **kw):
    """Sum over towers exactly like TFNN.optimSetup/sum_grads (TFModel.py:315-319,
    342-377): losses summed, gradients summed, lossVec concatenated."""
    outs = [loss_and_grad(theta, fd, **kw) for fd in tower_feeds]
    res = dict(loss=sum(o["loss"] for o in outs), BCloss=sum(o["BCloss"] for o in outs),
               ICloss=sum(o["ICloss"] for o in outs), varLoss=sum(o["varLoss"] for o in outs),
               lossVec=np.concatenate([o["lossVec"] for o in outs]),
               R=np.concatenate([o["R"] for o in outs]), Rabs=np.concatenate([o["Rabs"] for o in outs]),
               detJ=np.concatenate([o["detJ"] for o in outs]))
    res["grad"] = None if outs[0]["grad"] is None else sum(o["grad"] for o in outs)
    res["tower0"] = outs[0]
    return res


def strong_residual(theta, X, diff, vel, diff_dx, source, dim, inpDim, layerWidth, activation,
                    timeDependent, dtype=np.float64):
    """res = -u_t + kappa*Lap(u) - (vel - grad kappa).grad u + s   (TFModel.py:750-754).
    Second derivatives by forward-over-forward propagation of (z, dz, d2z)."""
    act = act_id(activation)
    c = lambda v: np.asarray(v, dtype=np.float32).astype(dtype)
    theta = c(theta); X = c(X).reshape(-1, inpDim)
    Ws, bs = unpack(theta, inpDim, layerWidth)
    L = len(layerWidth)
    P = X.shape[0]
    ndir = dim + (1 if timeDependent else 0)
    a = X
    for l in range(L):
        z = a @ Ws[l] + bs[l]
        if l == 0:
            dz = np.broadcast_to(Ws[0][:ndir, None, :], (ndir,) + z.shape)
            d2z = np.zeros((dim,) + z.shape, dtype=dtype)
        else:
            dz = da @ Ws[l]
            d2z = d2a @ Ws[l]
        a = _act(z, act)
        d1 = _d1(a, act)
        d2 = d1 * _d2_over_d1(a, act)
        da = d1[None] * dz
        d2a = d2[None] * dz[:dim] ** 2 + d1[None] * d2z
    u = a @ Ws[L][:, 0] + bs[L][0]
    du = da @ Ws[L][:, 0]                      # [ndir,P]
    lap = (d2a @ Ws[L][:, 0]).sum(axis=0)      # [P]
    res = np.zeros(P, dtype=dtype)
    if timeDependent:
        res = res - du[dim]
    res = res + c(diff).reshape(P) * lap
    vd = c(vel).reshape(P, dim) - c(diff_dx).reshape(P, dim)
    res = res - np.einsum("pk,kp->p", vd, du[:dim])
    res = res + c(source).reshape(P)
    return u, res


def lossvec_tolerance(res, rel=2e-6):
    """Conditioning-aware bound for the FP32 per-test-function field lossVec_i = detJ_i R_i^2.
    R_i is a cancelling sum (on the reference's tables the summed terms are ~1e3 x larger than
    R_i), so an FP32 evaluation — the reference's TF graph included — can only deliver R_i to
    about eps32 * sum|terms|.  Allowed: |dR_i| <= rel * Rabs_i (rel = 2e-6 ~ 34 ulp of the term
    magnitude), hence |d lossVec_i| <= detJ_i (2 |R_i| dR_i + dR_i^2)."""
    dR = rel * res["Rabs"]
    return res["detJ"] * (2.0 * np.abs(res["R"]) * dR + dR * dR)


def bout_tolerance(res, feed, timeDependent, rel=2e-6, base=1e-5):
    """Conditioning-aware bound for the gradient of the output bias.  g(b_out) contains
    sum_p ubar_p = -sum_i lambda_i sum_q w_q dNt_iq, a cancelling sum (on the reference's tables
    sum_q dNt_iq = 0 exactly), so an FP32 evaluation can only deliver it to eps32-level relative to
    sum_p |ubar_p|.  Allowed: base * |g| + rel * sum_p |ubar_p|  (ubar = -lambda dNt,
    lambda = 2 w2 detJ_i w_q R_i, App. A.3)."""
    if not timeDependent:
        return base * abs(float(res["grad"][-1]))
    nb, integNum = [int(v) for v in feed["intShape"]]
    dNt = np.abs(np.asarray(feed["dNt"], dtype=np.float64).reshape(nb, integNum))
    wq = np.ones((1, integNum)) if feed.get("integW") is None else np.abs(np.asarray(feed["integW"], dtype=np.float64).reshape(1, integNum))
    w2 = float(np.asarray(feed["w"], dtype=np.float64).reshape(3)[2])
    detJ = np.asarray(res["detJ"], dtype=np.float64).reshape(-1, 1) * np.ones((nb, 1))
    lam = 2.0 * w2 * detJ * np.abs(np.asarray(res["R"], dtype=np.float64).reshape(nb, 1))
    return base * abs(float(res["grad"][-1])) + rel * float(np.sum(lam * wq * dNt))


def varloss_tolerance(res, rel=2e-6, base=1e-5):
    """Bound for the summed variational loss on ill-conditioned (real) tables: the 1e-5 relative
    bar plus the worst-case accumulation of the per-test-function FP32 conditioning bound."""
    return base * abs(res["varLoss"]) + float(np.sum(lossvec_tolerance(res, rel)))


# ---------------------------------------------------------------- optimizers (TF 1.x)
def adam_step(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update (TF 1.10 adam.py `_apply_dense`):
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; theta -= lr_t*m/(sqrt(v)+eps).  t is the
    1-based step count.  (epsilon is added to sqrt(v), not to the bias-corrected sqrt.)"""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def rmsprop_step(theta, g, ms, mom, lr=1e-3, decay=0.9, momentum=0.0, eps=1e-10):
    """tf.train.RMSPropOptimizer (TF 1.10 defaults decay=.9, momentum=0, eps=1e-10;
    `ms` is initialised to ONES by TF): ms=decay*ms+(1-decay)g^2;
    mom=momentum*mom+lr*g/sqrt(ms+eps); theta-=mom."""
    ms = decay * ms + (1 - decay) * g * g
    mom = momentum * mom + lr * g / np.sqrt(ms + eps)
    return theta - mom, ms, mom


def loss_and_grad_chunked(theta, feed, chunk_tf=1024, workers=None, **kw):
    """`loss_and_grad` on a large feed, evaluated chunk by chunk over the test functions.

    The variational loss and its gradient are sums over test functions (TFModel.py:662-664) and the
    boundary/initial term does not depend on them, so the feed is cut into ranges of `chunk_tf` test
    functions: the first chunk carries the full loss weights, the others w = [0, 0, w2].  Chunks run on a
    thread pool (NumPy releases the GIL in BLAS and ufunc loops) with BLAS itself limited to one thread per
    call: OpenBLAS 0.3.30's pthreads pool gave wrong, run-to-run varying products when entered from eight
    Python threads at once (seen here: gradient off by up to 2.6 relative).  Same dict as `loss_and_grad`."""
    import concurrent.futures as cf
    import os
    from threadpoolctl import threadpool_limits
    nb, integNum = [int(v) for v in feed["intShape"]]
    w = np.asarray(feed["w"], dtype=np.float64).reshape(3)
    detJvec = bool(feed.get("detJvec", False))

    def piece(lo):
        hi = min(nb, lo + chunk_tf)
        f = dict(feed)
        for k in ("Input", "gcoef", "source", "N", "dNt"):
            v = feed.get(k)
            if isinstance(v, np.ndarray) and v.dtype != object and v.shape[0] == nb * integNum:
                f[k] = v[lo * integNum:hi * integNum]
        if detJvec:
            f["detJ"] = np.asarray(feed["detJ"]).reshape(-1)[lo:hi]
        f["intShape"] = [hi - lo, integNum]
        f["w"] = w if lo == 0 else np.array([0.0, 0.0, w[2]])
        return loss_and_grad(theta, f, **kw)

    starts = list(range(0, nb, chunk_tf))
    workers = workers or min(len(starts), os.cpu_count() or 1)
    if workers > 1:
        with threadpool_limits(limits=1), cf.ThreadPoolExecutor(max_workers=workers) as ex:
            outs = list(ex.map(piece, starts))
    else:
        outs = [piece(lo) for lo in starts]
    first = outs[0]
    varLoss = float(sum(o["varLoss"] for o in outs))
    res = dict(BCloss=first["BCloss"], ICloss=first["ICloss"], varLoss=varLoss,
               loss=w[0] * first["BCloss"] + w[1] * first["ICloss"] + w[2] * varLoss,
               lossVec=np.concatenate([o["lossVec"] for o in outs]), R=np.concatenate([o["R"] for o in outs]),
               Rabs=np.concatenate([o["Rabs"] for o in outs]), detJ=np.concatenate([o["detJ"] for o in outs]))
    res["grad"] = None if first["grad"] is None else np.sum([o["grad"] for o in outs], axis=0)
    return res
