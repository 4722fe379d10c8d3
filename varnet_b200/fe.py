"""Gauss-Legendre points and multilinear hat test functions on [-1,1]^d.

Host-side mirror of the reference's `FE` class (`/root/reference/FiniteElement.py:52-434`):
same constructor, same attribute names and the same numbers (validated against the
reference's own code in tests/test_tables_vs_reference.py and pinned by
tests/golden/fe_tables.npz).  Written as tensor-product array code instead of the
reference's per-dimension loops and recursion; multi-index order is "first coordinate
slowest" (`FiniteElement.py:143-151,179-188`).

Only the pieces the training hot path consumes are provided (`basisTot` and the tables
behind it); `massVec` / `testComp` (plots) are out of scope.
"""
import itertools

import numpy as np


class FE:
    """FE(dim, integPnum): compactly supported test functions and their quadrature."""

    def __init__(self, dim=2, integPnum=2):
        if integPnum > 3:
            raise ValueError('higher order integration needs code modification!')
        if integPnum == 2:                                    # FiniteElement.py:100-102
            integP = 1 / np.sqrt(3) * np.array([-1, 1])
            integW = np.ones(2)
        elif integPnum == 3:                                  # FiniteElement.py:103-105
            integP = np.sqrt(3 / 5) * np.array([-1, 0, 1])
            integW = 1 / 9 * np.array([5, 8, 5])
        else:
            raise ValueError('integPnum must be 2 or 3!')
        if dim not in (1, 2, 3):
            raise ValueError('FE dimension must be 1, 2 or 3!')
        self.dim = dim
        self.basisNum = 2 ** dim
        self.nodeNum = 3 ** dim
        self.integPnum = integPnum
        self.integP = integP
        self.integW = integW
        self.IntegPnum = integPnum ** dim
        # multi-indices of the 2^d corner bases, first coordinate slowest
        self.basMultiInd = np.array(list(itertools.product([-1, 1], repeat=dim)), dtype=int)
        # tensor grid of Gauss points, same ordering
        self.IntegP = np.array(list(itertools.product(integP, repeat=dim)), dtype=float).reshape(self.IntegPnum, dim)
        self.basVal = self._basis_values()
        self.basDeriVal = self._basis_derivatives()
        self.elemCoord = self._element_corners()
        self.delta = self._gauss_offsets()
        self.IntegW = self._weights()

    # factors 0.5*(1 + i_d*xi_d), multiplied from the last coordinate inwards so that the
    # floating-point association equals the reference recursion (FiniteElement.py:217-220)
    def _factor(self, d):
        ind = self.basMultiInd[:, d][:, None]                 # [2^d, 1]
        xi = self.IntegP[:, d][None, :]                       # [1, ip^d]
        return 0.5 * (1 + ind * xi)

    def _basis_values(self):
        out = self._factor(self.dim - 1)
        for d in range(self.dim - 2, -1, -1):
            out = self._factor(d) * out
        return out

    def _basis_derivatives(self):
        dv = np.zeros([self.dim, self.basisNum, self.IntegPnum])
        for k in range(self.dim):                             # differentiate w.r.t. coordinate k
            def term(d):
                if d == k:
                    return 0.5 * self.basMultiInd[:, d][:, None] * np.ones([1, self.IntegPnum])
                return self._factor(d)
            out = term(self.dim - 1)
            for d in range(self.dim - 2, -1, -1):
                out = term(d) * out
            dv[k] = out
        return dv

    def _element_corners(self):
        """elemCoord[d, e, c]: corner c of the element that carries basis e of the test function
        centred at the origin, in units of the element size (FiniteElement.py:298-324)."""
        order = self.basMultiInd.astype(float)                # [2^d, d]
        # 0.5*(corner_c - corner_e)
        return np.stack([0.5 * (order[None, :, d] - order[:, None, d]) for d in range(self.dim)], axis=0)

    def _gauss_offsets(self):
        delta = np.zeros([self.dim, self.basisNum, self.IntegPnum])
        for d in range(self.dim):                             # isoparametric map (FiniteElement.py:344-346)
            delta[d] = np.dot(self.elemCoord[d], self.basVal)
        return delta

    def _weights(self):
        if np.all(self.integW == 1.0):                        # trivial weights -> None (FiniteElement.py:369)
            return None
        w = self.integW
        W = w
        for _ in range(self.dim - 1):
            W = np.multiply.outer(W, w)
        W = W.reshape(1, self.IntegPnum)
        # association order of the reference: (w_i*w_j)*w_k
        return np.repeat(W, repeats=self.basisNum, axis=0)

    def basisTot(self, nt, hVec):
        """Tables for `nt` identical test functions of size hVec (FiniteElement.py:392-434).

        Returns integNum, nT, detJ, delta[dim,integNum], intWeight[1,integNum]|None,
        N[nT,1], dN[nT,dim] — N and dN are the per-test-function tables tiled nt times."""
        per = self.periodic_tables(hVec)
        integNum = per["integNum"]
        nT = nt * integNum
        N = np.tile(per["N"].reshape(integNum, 1), reps=[nt, 1])
        dN = np.tile(per["dN"], reps=[1, nt]).T
        return integNum, nT, per["detJ"], per["delta"], per["intWeight"], N, dN

    def periodic_tables(self, hVec):
        """The integNum-periodic part of `basisTot` (one test function), un-tiled.
        The device-side table generator consumes these directly."""
        hVec = np.asarray(hVec, dtype=float)
        integNum = self.basisNum * self.IntegPnum
        detJ = np.prod(0.5 * hVec)
        delta = np.reshape(self.delta, [self.dim, integNum])
        intWeight = None if self.IntegW is None else np.reshape(self.IntegW, [1, integNum])
        N = np.reshape(self.basVal, [integNum])
        nablaPhi = np.reshape(self.basDeriVal, [self.dim, integNum])
        dN = 2 / hVec.reshape(self.dim, 1) * nablaPhi         # [dim, integNum]
        return dict(integNum=integNum, detJ=detJ, delta=delta, intWeight=intWeight, N=N, dN=dN)
