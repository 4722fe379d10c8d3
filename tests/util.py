"""Shared helpers for the parity tests (synthetic feeds in the reference's feed-dict format)."""
import numpy as np

from oracle import graph_oracle as go


def synth_feed(rng, dim, inpDim, nb, integNum, nbi, bDof, timeDependent=True, isSource=False, integWflag=False,
               detJvec=False, w=(3.0, 5.0, 7.0)):
    """Random feed dict with the keys/shapes of VarNetUtility.py:840-854."""
    P = nb * integNum
    return dict(
        Input=rng.uniform(-1, 1, (P, inpDim)), gcoef=rng.randn(P, dim), source=rng.randn(P, 1), N=rng.rand(P, 1),
        dNt=rng.randn(P, 1) if timeDependent else [[None]],
        biInput=rng.uniform(-1, 1, (nbi, inpDim)), biLabel=rng.randn(nbi, 1), bDof=bDof, intShape=[nb, integNum],
        integW=rng.rand(1, integNum) + 0.5 if integWflag else None, biDimVal=2.0, detJvec=detJvec,
        detJ=(rng.rand(nb, 1) * 1e-2 + 1e-3) if detJvec else 1.3e-2, w=np.array(w, dtype=np.float64))


def make_engine(feed, dim, inpDim, layerWidth, activation, timeDependent, lossOpt, theta, device=0, optimizer="adam",
                dtype=None):
    from varnet_b200._capi import Engine
    eng = Engine(dim, inpDim, layerWidth, activation, timeDependent, lossOpt["isSource"], lossOpt["integWflag"],
                 optimizer=optimizer, device=device)
    eng.set_params(theta)
    eng.upload_points(feed["Input"], feed["gcoef"], feed["source"], feed["N"], feed["dNt"], feed["intShape"],
                      feed["integW"], feed["detJ"], bool(feed.get("detJvec", False)), dtype=dtype)
    eng.upload_bic(feed["biInput"], feed["biLabel"], feed["bDof"], feed["biDimVal"], dtype=dtype)
    eng.set_weights(feed["w"])
    return eng


def rel_inf(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


def layer_slices(inpDim, layerWidth):
    """[(name, slice)] of every Keras variable in the flat parameter vector."""
    out, off = [], 0
    for l, (i, o) in enumerate(go.layer_sizes(inpDim, layerWidth)):
        out.append(("kernel_%d" % l, slice(off, off + i * o))); off += i * o
        out.append(("bias_%d" % l, slice(off, off + o))); off += o
    return out
