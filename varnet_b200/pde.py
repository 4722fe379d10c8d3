"""Advection-diffusion problem definition (host side).

Mirror of the reference's `ADPDE` (`/root/reference/ADPDE.py:56-246`):

    dc/dt = div(diff grad c) - vel . grad c + s,     a grad c . n + b c = g on Gamma_i.

Same constructor signature, attribute names (`diffFun, velFun, sourceFun, d_diffFun, BCs,
BCtype, IC, cEx, MORvar, MORfunInd`, and `diff/vel/source` only when given as constants —
`VarNet` keys `lossOpt['isSource']` off `hasattr(PDE,'source')`, VarNet.py:184-185).
Plot helpers are out of scope.
"""
import numpy as np

from .hostutil import is_number


def _const_field(value, ncols):
    def field(x, t=0):
        return value * np.ones([np.shape(x)[0], ncols])
    return field


class ADPDE:
    def __init__(self, domain, diff, vel, source=0.0, timeDependent=False, tInterval=None, BCs=None, IC=None,
                 cEx=None, MORvar=None, d_diff=None):
        timeDependent = tInterval is not None                 # the flag is derived, as in the reference
        for name, val in (("diffusivity field", diff), ("velocity field", vel)):
            if not is_number(val) and not callable(val):
                raise ValueError(name + ' must be constant or callable!')
        if not is_number(source) and not callable(source):
            raise ValueError('source function must be constant or callable!')
        if BCs is not None and type(BCs) is not list:
            raise ValueError('BCs must be empty or a list of [a, b, g(x,t)]!')
        if BCs is not None and len(BCs) != domain.bIndNum:
            raise ValueError('number of BCs does not match number of boundaries in domain!')
        if timeDependent and IC is None:
            raise ValueError('initial condition must be provided for time-dependent problems!')
        if cEx is not None and not callable(cEx):
            raise ValueError('exact solution must be a callable function!')
        if d_diff is not None and not is_number(d_diff) and not callable(d_diff):
            raise ValueError('diffusivity gradient must be constant or callable!')
        dim = domain.dim

        if callable(diff):
            self.diffFun = diff
        else:
            self.diff, self.diffFun = diff, _const_field(diff, 1)
        if callable(vel):
            self.velFun = vel
        else:
            self.vel, self.velFun = vel, _const_field(np.asarray(vel, dtype=float) if np.size(vel) > 1 else vel, dim)
        if callable(source):
            self.sourceFun = source
        else:
            self.source, self.sourceFun = source, _const_field(source, 1)
        if callable(d_diff):
            self.d_diffFun = d_diff
        else:
            self.d_diff = 0.0 if d_diff is None else d_diff
            self.d_diffFun = _const_field(self.d_diff, dim)

        # boundary conditions in the standard [a, b, g] form; default = homogeneous Dirichlet
        nB = domain.bIndNum
        BCs = [[] for _ in range(nB)] if BCs is None else list(BCs)
        for k in range(nB):
            bc = BCs[k]
            if isinstance(bc, (list, tuple)) and len(bc) == 0:
                BCs[k] = [0.0, 1.0, lambda x, t=0: np.zeros([len(x), 1])]
            elif len(bc) != 3:
                raise ValueError('BCs must be specified as a list of [a, b, g(x,t)]!')
            elif not callable(bc[2]):
                BCs[k] = [bc[0], bc[1], _const_field(bc[2], 1)]
        BCtype = ['Dirichlet' if bc[0] == 0 else ('Neumann' if bc[1] == 0 else 'Robin') for bc in BCs]

        if timeDependent and not callable(IC):
            IC = (lambda x, t=0: np.zeros([len(x), 1])) if (isinstance(IC, (list, tuple)) and len(IC) == 0) \
                else _const_field(IC, 1)

        if MORvar is not None:
            self.MORfunInd = self._mor_lookup(MORvar, BCs, IC, d_diff)
        self.dim = dim
        self.domain = domain
        self.timeDependent = timeDependent
        self.tInterval = tInterval
        self.BCs = BCs
        self.BCtype = BCtype
        self.IC = IC
        self.cEx = cEx
        self.MORvar = MORvar

    def _mor_lookup(self, MORvar, BCs, IC, d_diff):
        """Which PDE field each parametric function feeds (ADPDE.py:201-236)."""
        table = {'diff': None, 'vel': None, 'source': None, 'IC': None, 'd_diff': None}
        fields = {'diff': self.diffFun, 'vel': self.velFun, 'source': self.sourceFun, 'IC': IC,
                  'd_diff': self.d_diffFun}
        BCind = [None] * len(BCs)
        for i, fn in enumerate(MORvar.funcHandles):
            hit = next((k for k, f in fields.items() if fn == f), None)
            if hit is not None:
                table[hit] = i
                continue
            for b, bc in enumerate(BCs):
                if fn == bc[2]:
                    BCind[b] = i
        if table['diff'] is not None and callable(d_diff) and table['d_diff'] is None:
            raise ValueError('\'diff\' has extra input arguments but \'d_diff\' does not!')
        table['BCs'] = BCind
        table['inpData'] = True if any(table[k] is not None for k in ('diff', 'vel', 'source')) else None
        table['biData'] = True if (any(b is not None for b in BCind) or table['IC'] is not None) else None
        return table
