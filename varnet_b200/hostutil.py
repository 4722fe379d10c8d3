"""Small host helpers shared by the table builders (pure NumPy).

Functional equivalents of the handful of `UtilityFunc.UF` methods the hot path's host
side relies on (`/root/reference/UtilityFunc.py`: isnone :113, isempty :101, vstack :158,
pairMats :301-339, l2Err :485, buildDict :470)."""
import numbers

import numpy as np


def is_empty(x):
    if isinstance(x, (list, dict, tuple)):
        return len(x) == 0
    return x is not None and np.size(x) == 0


def is_none(x):
    """True when x is None or an object array / list holding a None (e.g. dNt=[[None]])."""
    if x is None:
        return True
    if isinstance(x, (list, tuple)):
        return any(is_none(v) for v in x)
    if isinstance(x, np.ndarray) and x.dtype == object:
        return any(v is None for v in x.ravel())
    return False


def is_number(x):
    if isinstance(x, (list, tuple)):
        return all(is_number(v) for v in x)
    if isinstance(x, np.ndarray):
        return x.dtype.kind in "fiub"
    return isinstance(x, numbers.Number)


def stack_rows(blocks):
    """vstack that skips empty blocks and returns [] when nothing is left."""
    keep = [b for b in blocks if not is_empty(b)]
    return np.vstack(keep) if keep else []


def pair_rows(a, b):
    """Cartesian pairing used for space x time grids: every row of `a` is repeated for each
    row of `b` (rows of `a` slow, rows of `b` fast) and the columns are concatenated."""
    if is_empty(a):
        return b
    if is_empty(b):
        return a
    a = np.asarray(a); b = np.asarray(b)
    return np.hstack([np.repeat(a, len(b), axis=0), np.tile(b, reps=[len(a), 1])])


def l2_err(x_true, x_app):
    x_true = np.asarray(x_true, dtype=float).reshape(-1, 1)
    x_app = np.asarray(x_app, dtype=float).reshape(-1, 1)
    if x_true.size != x_app.size:
        raise ValueError('\'xTrue\' and \'xApp\' must have the same shape!')
    return np.linalg.norm(x_true - x_app) / np.linalg.norm(x_true)


def split_rows(vec, counts):
    """Consecutive row blocks of the given sizes; a trailing remainder becomes one more block
    (what UtilityFunc.listSegment does without a callback, UtilityFunc.py:407-455)."""
    if counts is None:
        return [vec]
    counts = [counts] if isinstance(counts, numbers.Number) else list(counts)
    out, start = [], 0
    for c in counts:
        out.append(vec[start:start + c])
        start += c
    if start < len(vec):
        out.append(vec[start:])
    return out


def rejection_sampling(func, smpfun, dof, dofT=None):
    """Draw `dof[i]` samples per segment with acceptance probability func(x)/max(func) (segment-wise max
    over the default grid `func()`), the scheme of UtilityFunc.rejectionSampling (UtilityFunc.py:342-404).
    The order of random draws (candidate batch, then one uniform vector per segment) is kept so that a
    seeded run selects the same points as the reference."""
    if isinstance(dof, numbers.Number):
        if dofT is not None:
            raise ValueError('\'dofT\' must be None for scalar \'dof\'')
        dof = [dof]
    m = len(dof)
    if m > 1 and dofT is None:
        raise ValueError('\'dofT\' must be provided when \'dof\' is a list!')
    fmax = [np.max(seg) for seg in split_rows(func(), dofT)]
    kept = [[] for _ in range(m)]
    count = [0] * m
    while any(count[i] < dof[i] for i in range(m)):
        samples = smpfun()
        cand = split_rows(samples, dofT)
        vals = split_rows(func(samples), dofT)
        for i in range(m):
            v = vals[i]
            accept = (np.random.uniform(size=[len(v), 1]) < (v / fmax[i])).reshape(len(v))
            kept[i] = stack_rows([kept[i], cand[i][accept]])
            count[i] += int(np.sum(accept))
    return np.vstack([kept[i][:dof[i], :] for i in range(m)])
