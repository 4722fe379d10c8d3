"""Parametric ("MOR") inputs: bookkeeping of the extra MLP inputs (host side).

Mirror of the reference's `MOR` (`/root/reference/MOR.py:52-254`): validates which trailing
arguments of the user's field functions are parameters, discretises their ranges and
enumerates the parameter batches.  `POD` (classical MOR comparison, unused by VarNet) is
out of scope.
"""
import numpy as np

from .hostutil import pair_rows


def _reorder(seq, order):
    return [seq[i] for i in order]


class MOR:
    def __init__(self, funcHandles, ArgNames, ArgRange):
        if type(funcHandles) is not list:
            if not callable(funcHandles):
                raise ValueError('\'funcHandles\' must be a list of callable functions!')
            funcHandles, ArgNames, ArgRange = [funcHandles], [ArgNames], [ArgRange]
        varNum, argInd, sortInd = [], [], []
        for i, func in enumerate(funcHandles):
            if not callable(func):
                raise ValueError('entries must be callable functions!')
            code = func.__code__
            params = code.co_varnames[:code.co_argcount]
            if type(ArgNames[i]) is not list:
                ArgNames[i], ArgRange[i] = [ArgNames[i]], [ArgRange[i]]
            pos = []
            for name in ArgNames[i]:
                if name not in params:
                    raise ValueError(name + ' is not an argument of ' + code.co_name + '!')
                pos.append(params.index(name))
            order = np.argsort(pos)
            pos = _reorder(pos, order)
            ArgNames[i] = _reorder(ArgNames[i], order)
            # parameters must be the contiguous tail of the signature (MOR.py:116-118)
            if not (pos[-1] == code.co_argcount - 1 and len(pos) == pos[-1] - pos[0] + 1):
                raise ValueError('variable arguments of ' + code.co_name +
                                 ' must be ordered and the last arguments to the function')
            if np.shape(ArgRange[i])[1] != 2:
                raise ValueError('dimension of the variable ranges for function ' + code.co_name +
                                 'are not equal to 2!')
            if len(ArgRange[i]) != len(pos):
                raise ValueError('number of variable ranges for function ' + code.co_name +
                                 'does not match the number of variable arguments!')
            ArgRange[i] = _reorder(ArgRange[i], order)
            varNum.append(len(pos)); argInd.append(pos); sortInd.append(order)
        self.funNum = len(funcHandles)
        self.funcHandles = funcHandles
        self.ArgNames = ArgNames
        self.ArgRange = ArgRange
        self.varNum = varNum
        self.argInd = argInd
        self.sortInd = sortInd

    def discretizeArg(self, discScheme, randFlag=False):
        """Per function: matrix of all parameter combinations, one per row (MOR.py:147-233)."""
        if type(discScheme) is not list:
            if not callable(discScheme):
                raise ValueError('\'discScheme\' must be a list!')
            discScheme = [discScheme]
        out = []
        for i, func in enumerate(self.funcHandles):
            name = func.__code__.co_name
            scheme = discScheme[i]
            if callable(scheme):
                grid = scheme()
                if np.shape(grid)[1] != self.varNum[i]:
                    raise ValueError('output dimension of the function handle to discretize ' + name +
                                     ' is not equal to its number of variable arguments!')
                out.append(grid)
                continue
            if np.size(scheme) not in (1, self.varNum[i]):
                raise ValueError('number of discretization numbers for function ' + name +
                                 ' does not match the number of variable arguments!')
            counts = np.tile(scheme, self.varNum[i]) if np.size(scheme) == 1 else _reorder(list(scheme), self.sortInd[i])
            grid = []
            for j in range(self.varNum[i]):
                lo, hi = self.ArgRange[i][j][0], self.ArgRange[i][j][1]
                n = int(counts[j])
                vals = np.sort(np.random.uniform(lo, hi, n)) if randFlag else np.linspace(lo, hi, n)
                grid = pair_rows(grid, vals.reshape(n, 1))
            out.append(grid)
        return out

    def argIndex(self, discArg):
        """Index combinations across functions: one row per MOR batch (MOR.py:236-254)."""
        idx = []
        for i in range(self.funNum):
            n = len(discArg[i])
            idx = pair_rows(idx, np.arange(n).reshape(n, 1))
        return idx
