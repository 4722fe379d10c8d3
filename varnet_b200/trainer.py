"""`VarNet`: trainer/orchestrator with the reference's API, driving the B200 backend.

Mirror of the reference's `VarNet` class for the MLP / weak-form path
(`/root/reference/VarNet.py`): constructor (`:70-203`), `timeDisc` (`:297-338`),
`trainingPoints` (`:504-600`), `biTrainPoints`/`biTrainData` (`:604-722`), `PDEinpData`
(`:726-774`), `trainData` (`:778-897`), `MORargExtract` (`:901-1049`), `splitLoss` (`:1053-1090`),
`trainWeight` (`:1094-1146`), `train` (`:1197-1421`), `evaluate` (`:1510-1595`), `residual`
(`:1599-1692`), residual-driven resampling `optTrainPoints`/`optBiTrainPoints` (`:1696-1966`),
`loadModel` (checkpoint = flat vector, `:1426-1506`).

Out of scope here: the RNN/GRU variant, matplotlib plotting (`simRes`, the residual
contour plots inside `optTrainPoints`), the periodic `weightUpdate` heuristic.
"""
import glob
import math
import os
import time
import warnings

import numpy as np

from .backend import TFNN, GlobalInit
from .hostutil import (is_empty, is_none, is_number, l2_err, pair_rows, rejection_sampling, split_rows,
                       stack_rows)
from .tables import FIXData, ManageTrainData, TableView


def _rank():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
    except Exception:
        pass
    return 0


def sync_numpy_rng():
    """Under torchrun every rank builds the same tables and slices its own tower out of them, so the `np.random`
    draws of the 'random' / 'optimal' sampling schemes must agree across ranks: rank 0 draws a seed from its own
    generator, broadcasts it, and every rank re-seeds with it.  No-op in a single process."""
    try:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
    except Exception:
        return
    box = [int(np.random.randint(0, 2 ** 31 - 1))]
    dist.broadcast_object_list(box, src=0)
    np.random.seed(box[0])


class TrainLog:
    """Minimal text logger standing in for `TrainResult` (VarNetUtility.py:1149-1757).  Under torchrun only
    rank 0 writes the case file (the ranks share `folderpath`)."""

    def __init__(self, folderpath, verbose=True, saveFreq=100):
        self._writer = _rank() == 0
        verbose = verbose and self._writer
        self.folderpath, self.verbose, self.saveFreq = folderpath, verbose, saveFreq
        self.loss, self.lossComp, self.residual, self.error, self.inpIter = [], [], [], [], []
        self.trainWeight = None
        self.epoch_time = 0.0
        os.makedirs(folderpath, exist_ok=True)
        self._path = os.path.join(folderpath, 'caseData.txt')

    def writeCase(self, string):
        if not self._writer:
            return
        with open(self._path, 'a') as f:
            f.write(string + '\n')

    writeComment = writeCase

    def initializeCase(self, vn, argDict):
        tf = vn.tfData
        lines = ['VarNet (B200 backend) case file', 'dim: %d  inpDim: %d  layerWidth: %s  activation: %s' %
                 (vn.dim, tf.inpDim, tf.layerWidth, tf.activationFun[0]),
                 'parameters: %d  processors: %s  optimizer: %s  learning rate: %g' %
                 (tf.model.count_params(), tf.processors, tf.optimizer_name, tf.learning_rate),
                 'training arguments: %s' % {k: v for k, v in argDict.items() if k != 'self'}]
        if not self._writer:
            return
        with open(self._path, 'w') as f:
            f.write('\n'.join(lines) + '\n')

    def iterOutput(self, epoch, current_loss, min_loss, epoch_time, resVal, err, lossComp, lossVec):
        self.loss.append(current_loss)
        self.epoch_time = epoch_time
        if epoch % self.saveFreq == 0:
            self.residual.append(resVal); self.error.append(err)
            if lossComp is not None:
                self.lossComp.append(lossComp)
            line = 'epoch %d  loss %.6g  best %.6g  residual %s  error %s  avg iter time %.3g s' % (
                epoch, current_loss, min_loss, resVal, err, epoch_time / epoch)
            self.writeCase(line)
            if self.verbose:
                print(line)


class VarNet:
    def __init__(self, PDE, layerWidth=[20], modelId='MLP', activationFun=None, discNum=20, bDiscNum=[],
                 tDiscNum=[], MORdiscScheme=None, processors=None, controller=None, integPnum=2,
                 optimizer='adam', learning_rate=0.001, seed=None):
        dim, td, MORvar = PDE.dim, PDE.timeDependent, PDE.MORvar
        if np.size(discNum) != 1 and np.size(discNum) != dim:
            raise ValueError('dimension of the number of discretizations does not match dimension of the domain!')
        if np.size(discNum) == 1:
            discNum = [discNum] * dim if np.shape(discNum) == () else [discNum[0]] * dim
        if np.size(bDiscNum) != 1:
            raise ValueError('density of boundary discretizations must be a scalar!')
        if modelId != 'MLP':
            raise ValueError('only modelId=\'MLP\' is supported by the B200 backend')
        if td and is_empty(tDiscNum):
            raise ValueError('time discretization number must be provided for time-dependent PDEs!')
        if activationFun is None:
            activationFun = 'sigmoid'                                   # VarNet.py:162-164
        if type(layerWidth) is not list:
            raise ValueError('layer widths should be given in a list!')
        if MORvar is not None and MORdiscScheme is None:
            raise ValueError('\'MORdiscScheme\' must be given for MOR!')
        inpDim = dim + (1 if td else 0) + (sum(MORvar.varNum) if MORvar is not None else 0)
        lossOpt = {'integWflag': integPnum != 2,
                   'isSource': not (hasattr(PDE, 'source') and PDE.source == 0.0)}     # VarNet.py:182-185
        self.dim, self.discNum, self.bDiscNum, self.tDiscNum = dim, discNum, bDiscNum, tDiscNum
        self.MORdiscScheme, self.modelId, self.PDE = MORdiscScheme, modelId, PDE
        self.fixData = FIXData(self, integPnum)
        self.fixData.setInputData(self)
        self.tfData = TFNN(dim, inpDim, layerWidth, modelId, activationFun, td, None, processors, controller,
                           lossOpt, optimizer, learning_rate, seed=seed)

    # ------------------------------------------------------------------ discretisation
    def timeDisc(self, tdof=None, rfrac=0, sortflg=True, discTol=None):
        PDE = self.PDE
        if not PDE.timeDependent:
            raise Exception('The problem is time-independent!')
        rfrac = min(max(rfrac, 0), 1)
        tdof = self.tDiscNum if tdof is None else tdof
        t0, t1 = PDE.tInterval[0], PDE.tInterval[1]
        ht = (t1 - t0) / tdof
        tol = ht if discTol is None else np.asarray(discTol).item()
        nrand = math.floor(tdof * rfrac)
        parts = [np.random.uniform(t0 + tol, t1, nrand)] if nrand else []
        parts.append(np.linspace(t0 + tol, t1, tdof - nrand))          # last centre sits on t1 (App. C.3)
        t = np.hstack(parts)
        if rfrac > 0 and sortflg:
            t = np.sort(t)
        return ht, t.reshape(tdof, 1)

    def trainingPoints(self, smpScheme='uniform', frac=0.5, addTrainPts=True, suppFactor=1.0):
        """Quadrature-point coordinates Input[nT, feDim]: centre + h*delta, space index slow, time
        index fast, Gauss index fastest (VarNet.py:576-586)."""
        if smpScheme != 'uniform':
            sync_numpy_rng()                  # identical draws on every rank (each slices its tower from the same tables)
        if smpScheme == 'optimal':
            return self.optTrainPoints(frac, addTrainPts, suppFactor)
        rfrac = frac if smpScheme == 'random' else 0.
        fd, PDE, dim = self.fixData, self.PDE, self.dim
        td, domain = PDE.timeDependent, PDE.domain
        if td:
            tDiscNum = self.tDiscNum
            ht, t_coord = self.timeDisc(rfrac=rfrac)
        else:
            tDiscNum, t_coord = 1, []
        mesh = domain.getMesh(self.discNum, self.bDiscNum, rfrac=rfrac)
        if smpScheme == 'random' and mesh.dof < fd.dof:
            coord = mesh.coordinates
            while len(coord) < fd.dof:
                coord = np.vstack([coord, domain.getMesh(self.discNum, self.bDiscNum, rfrac=1.).coordinates])
            mesh.dof, mesh.coordinates = fd.dof, coord[:fd.dof, :]
        he, coord, delta = mesh.he, mesh.coordinates, fd.delta
        nt, nT = fd.nt, fd.nT
        cols = []
        for d in range(dim):
            c = np.repeat(coord[:, d], repeats=tDiscNum).reshape(nt, 1) + he[d] * delta[d, :]
            cols.append(c.reshape(nT))
        if td:
            tc = np.tile(t_coord, reps=[fd.dof, 1]) + ht * delta[-1, :]
            cols.append(tc.reshape(nT))
        Input = np.stack(cols, axis=1)
        # what a backend needs to rebuild this table on the device (only valid for the plain uniform mesh)
        self._uniform_mesh = dict(Input=Input, coord=coord, t_coord=(t_coord if td else None)) if smpScheme == 'uniform' else None
        biInput, biDof = self.biTrainPoints(mesh, t_coord)
        return Input, [], biInput, biDof

    def _gen_spec(self, Input, diff, vel, src):
        """Recipe for building this batch's point table on the device (vn_generate_table_f64) instead of uploading
        nT-row arrays: valid when `Input` is the uniform-mesh table of `trainingPoints()` and the PDE coefficients are
        constants over the batch (constant kappa / velocity, or a MOR batch whose parameter is one value).  The
        device formulas and operation order are those of `trainingPoints` / `trainData` (VarNet.py:576-586,837), so the
        generated table is bit-identical to the uploaded one (tests/test_gpu_parity.py)."""
        um = getattr(self, "_uniform_mesh", None)
        fd = self.fixData
        if um is None or um["Input"] is not Input or fd.detJvec or self.tfData.lossOpt['isSource']:
            return None
        const = lambda a: a is not None and np.size(a) > 0 and float(np.ptp(np.asarray(a, dtype=float), axis=0).max()) == 0.0
        if not (const(diff) and const(vel)):
            return None
        q = fd.integNum
        dN = np.hstack([fd.dNx[:q], fd.dNt[:q]]) if fd.timeDependent else fd.dNx[:q]
        return dict(coord=um["coord"], tcoord=um["t_coord"], hVec=fd.hVec, delta=fd.delta, N=fd.N[:q], dN=dN,
                    diff=float(np.asarray(diff, dtype=float).reshape(-1)[0]),
                    vel=np.asarray(vel, dtype=float).reshape(-1, self.dim)[0].copy(), source=0.0,
                    nb=fd.nt, integNum=q, integW=fd.integW, detJ=fd.detJ)

    # ------------------------------------------------------------------ residual-driven sampling
    def optTrainPoints(self, frac=0.25, addTrainPts=True, suppFactor=1.0):
        """Test-function centres drawn by rejection sampling on |strong-form residual| (VarNet.py:1696-1860),
        stacked before the (possibly thinned) uniform ones; with suppFactor != 1 the new supports are
        scaled and detJ becomes a per-test-function vector (FIXData.updateOptimData).  Plots are skipped."""
        fd, PDE, dim = self.fixData, self.PDE, self.dim
        td, domain = PDE.timeDependent, PDE.domain
        nt, nT, hVec, delta, integNum = fd.nt0, fd.nT, fd.hVec, fd.delta, fd.integNum
        keep = 1 if addTrainPts else (1 - frac) ** (1 / fd.feDim)
        if td:
            tDisc2 = math.ceil(keep * self.tDiscNum)
            _, t_coord = self.timeDisc(tDisc2)
        else:
            tDisc2, t_coord = 1, []
        mesh = domain.getMesh([math.ceil(keep * d) for d in self.discNum], self.bDiscNum)
        uniform_part = pair_rows(mesh.coordinates, t_coord)
        nt1 = math.ceil(frac * nt) if addTrainPts else nt - mesh.dof * tDisc2
        scaled = np.abs(suppFactor - 1.0) >= 1.e-15
        tole = suppFactor * hVec[:dim] if scaled else None
        tolt = suppFactor * hVec[-1] if (scaled and td) else None

        def resfun(pts=None):
            return np.abs(self.residual(pts)[1])

        def smpfun():
            tc = self.timeDisc(rfrac=1, sortflg=False, discTol=tolt)[1] if td else []
            m = domain.getMesh(self.discNum, self.bDiscNum, rfrac=1, sortflg=False, discTol=tole)
            return pair_rows(m.coordinates, tc)

        optimal_part = rejection_sampling(resfun, smpfun, nt1)
        if addTrainPts:
            nt = nt1 + nt
            nT = nt * integNum
        pts = np.vstack([optimal_part, uniform_part])
        coord = pts[:, :dim]
        if td:
            t_coord = pts[:, dim:dim + 1]
            if not scaled:                                      # identical supports: order by time
                order = np.argsort(t_coord, axis=0).reshape(nt)
                coord, t_coord = coord[order], t_coord[order]
        biInput, biDof, _, _ = self.optBiTrainPoints(frac, addTrainPts)
        if scaled:
            supp = np.ones([nt, 1]); supp[:nt1, :] = suppFactor
        else:
            supp = 1.0
        cols = [(coord[:, d].reshape(nt, 1) + hVec[d] * delta[d, :] * supp).reshape(nT) for d in range(dim)]
        if td:
            cols.append((t_coord + hVec[-1] * delta[-1, :] * supp).reshape(nT))
        Input = np.stack(cols, axis=1)
        if addTrainPts:
            self.fixData.updateOptimData(frac, suppFactor)
        return Input, [], biInput, biDof

    def optBiTrainPoints(self, frac=0.25, addTrainPts=True):
        """Boundary/initial rows by rejection sampling on (model - label)^2 per boundary segment
        (VarNet.py:1864-1966)."""
        fd, PDE, dim = self.fixData, self.PDE, self.dim
        td, domain, tf = PDE.timeDependent, PDE.domain, self.tfData
        biDof = fd.biDof0
        keep = 1 if addTrainPts else (1 - frac) ** (1 / (fd.feDim - 1))
        t_coord = self.timeDisc(math.ceil(keep * self.tDiscNum))[1] if td else []
        bDisc2 = math.ceil(keep * self.bDiscNum) if self.bDiscNum is not None else None
        mesh = domain.getMesh([math.ceil(keep * d) for d in self.discNum], bDisc2)
        uniform_part, biDof2 = self.biTrainPoints(mesh, t_coord)
        biDof1 = [math.ceil(frac * b) for b in biDof] if addTrainPts else list(np.array(biDof) - np.array(biDof2))
        tw = next(t for t in tf.compTowers if t.local)          # weights are replicated: every rank evaluates locally

        def resfun(rows=None):
            rows = fd.uniform_biInput if rows is None else rows
            val = tf.sess.run(tf.model(tw.Input), {tw.Input: rows})
            return (val - self.biTrainData(rows, biDof)) ** 2

        def smpfun():
            tc = self.timeDisc(rfrac=1, sortflg=False)[1] if td else []
            m = domain.getMesh(self.discNum, self.bDiscNum, rfrac=1, sortflg=False)
            return self.biTrainPoints(m, tc)[0]

        optimal_part = rejection_sampling(resfun, smpfun, biDof1, biDof)
        biDofNew = list(np.array(biDof1) + np.array(biDof2)) if addTrainPts else biDof
        blocks = []
        for o, u, n in zip(split_rows(optimal_part, biDof1), split_rows(uniform_part, biDof2), biDofNew):
            rows = stack_rows([o, u])
            if td:
                order = np.argsort(rows[:, dim:dim + 1], axis=0).reshape(n)
                rows = np.hstack([rows[:, :dim][order], rows[:, dim:dim + 1][order]])
            blocks.append(rows)
        return np.vstack(blocks), biDofNew, optimal_part, uniform_part

    def biTrainPoints(self, mesh, t_coord):
        """Dirichlet boundary rows (space x time) per boundary, then initial-condition rows [x, 0]."""
        PDE = self.PDE
        bInput, biDof = [], []
        for b in range(PDE.domain.bIndNum):
            if PDE.BCtype[b] == 'Dirichlet':
                rows = pair_rows(mesh.bCoordinates[b], t_coord)
                biDof.append(len(rows)); bInput.append(rows)
        iInput = []
        if PDE.timeDependent:
            iInput = np.concatenate([mesh.coordinates, np.zeros([mesh.dof, 1])], axis=1)
            biDof.append(mesh.dof)
        return stack_rows([stack_rows(bInput), iInput]), biDof

    def biTrainData(self, biInput, biDof, biArg=[], biLabel0=[]):
        """Labels g/beta on Dirichlet rows and IC(x) on initial rows (VarNet.py:649-722).  biArg
        entries: {} = evaluate, dict = evaluate with MOR kwargs, None = reuse biLabel0."""
        PDE, dim = self.PDE, self.dim
        td, nB = PDE.timeDependent, PDE.domain.bIndNum
        if is_empty(biArg):
            biArg = [{} for _ in range(nB + (1 if td else 0))]
        labels, start, end = [], 0, 0
        for b in range(nB):
            end = start + biDof[b] if b < len(biDof) else start
            if PDE.BCtype[b] != 'Dirichlet':
                continue
            if biArg[b] is None:
                if is_empty(biLabel0):
                    raise ValueError('\'biLabel0\' must be provided for \'None\' arguments!')
                labels.append(biLabel0[start:end, :])
                continue                                               # (the reference does not advance here)
            beta, g = PDE.BCs[b][1], PDE.BCs[b][2]
            targs = [biInput[start:end, dim][np.newaxis].T] if td else []
            labels.append(g(biInput[start:end, :dim], *targs, **biArg[b]) / beta)
            start = end
        out = [stack_rows(labels)]
        if td:
            if biArg[-1] is None:
                if is_empty(biLabel0):
                    raise ValueError('\'biLabel0\' must be provided for \'None\' arguments!')
                out.append(biLabel0[end:, :])
            else:
                out.append(PDE.IC(biInput[end:, :dim], **biArg[-1]))
        return stack_rows(out)

    def PDEinpData(self, Input, inpArg=[]):
        """kappa, vel, s at the given space-time points (VarNet.py:726-774)."""
        PDE, dim = self.PDE, self.dim
        args = [{}, {}, {}] if is_empty(inpArg) else inpArg
        targs = [Input[:, -1][np.newaxis].T] if PDE.timeDependent else []
        funs = (PDE.diffFun, PDE.velFun, PDE.sourceFun)
        return tuple(None if a is None else f(Input[:, 0:dim], *targs, **a) for f, a in zip(funs, args))

    # ------------------------------------------------------------------ MOR arguments
    def MORargExtract(self, batch, MORdiscArg, defArg=None):
        """kwargs of the parametric functions for MOR batch `batch`, and the extra NN inputs.
        A function whose parameters did not change since the previous batch gets `defArg`
        (None = keep stored data) (VarNet.py:901-1049)."""
        PDE = self.PDE
        funInd, names = PDE.MORfunInd, PDE.MORvar.ArgNames
        argInd = self.fixData.MORargInd
        inpNN = []

        def pick(find, first_always):
            vals = MORdiscArg[find][argInd[batch, find], :]
            if batch > 0 or not first_always:
                if batch > 0:
                    prev = MORdiscArg[find][argInd[batch - 1, find], :]
                    if l2_err(vals, prev) < 1.e-10:
                        return None
            inpNN.extend(vals)
            return dict(zip(names[find], vals))

        biArg = []
        if funInd['biData']:
            for b in range(PDE.domain.bIndNum):
                f = funInd['BCs'][b]
                biArg.append(None if (PDE.BCtype[b] != 'Dirichlet' or f is None) else pick(f, False))
            if funInd['IC'] is not None:
                biArg.append(pick(funInd['IC'], True))
        inpArg = []
        if funInd['inpData']:
            for key in ('diff', 'vel', 'source'):
                f = funInd[key]
                if f is None:
                    inpArg.append(defArg)
                else:
                    got = pick(f, True)
                    inpArg.append(defArg if got is None else got)
        return biArg, inpArg, np.reshape(inpNN, [1, np.size(inpNN)])

    # ------------------------------------------------------------------ per-batch data
    def trainData(self, batch, MORdiscArg, tData, resCalc=False):
        """(Re)build only what changed: gcoef = kappa*dNx + vel*N (VarNet.py:837), the MOR input
        columns, BC/IC labels; create the feed dicts on first use (VarNet.py:778-897)."""
        if not tData.inputUpdated and MORdiscArg is None:
            return tData
        if tData.MORdataSaved:
            tData.loadMORData(batch)
            return tData
        fd = self.fixData
        Input, biInput, _, _, biLabel0, _, _, diff0, vel0 = tData.getAllData()
        nT = np.shape(Input)[0] if resCalc else fd.nT
        fresh = tData.inputUpdated
        if fresh:
            if batch != 0 and not resCalc:
                raise ValueError('\'batch\' variable must be reset when the space-time discretization is updated!')
            if MORdiscArg is None:
                biArg, inpArg, MORinpNN = [], [], None
            else:
                biArg, inpArg, MORinpNN = self.MORargExtract(batch, MORdiscArg, defArg={})
            biLabel = [] if resCalc else self.biTrainData(biInput, fd.biDof, biArg)
            if resCalc and MORdiscArg is None and np.shape(fd.uniform_input)[0] == nT:
                diff, vel, src = fd.uniform_inpData
            else:
                diff, vel, src = self.PDEinpData(Input, inpArg)
            gcoef = [] if resCalc else diff * fd.dNx + vel * fd.N
        else:
            biArg, inpArg, MORinpNN = self.MORargExtract(batch, MORdiscArg)
            funInd = self.PDE.MORfunInd
            if resCalc:
                biLabel = []
            else:
                biLabel = self.biTrainData(biInput, fd.biDof, biArg, biLabel0) if funInd['biData'] else None
            if funInd['inpData']:
                diff, vel, src = self.PDEinpData(Input, inpArg)
                if resCalc:
                    gcoef = []
                elif diff is None and vel is None:
                    gcoef = None
                else:
                    gcoef = (diff0 if diff is None else diff) * fd.dNx + (vel0 if vel is None else vel) * fd.N
            else:
                gcoef = src = diff = vel = None
        if MORinpNN is not None and not resCalc and getattr(self.tfData, 'supports_table_views', False):
            tData._views = True
            InpuTot = TableView(Input, None, 1, MORinpNN)                   # MOR columns stay per-batch constants
            biInpuTot = np.hstack([biInput, np.tile(MORinpNN, reps=[int(np.sum(fd.biDof)), 1])])
        elif MORinpNN is not None:
            InpuTot = np.hstack([Input, np.tile(MORinpNN, reps=[nT, 1])])
            biInpuTot = [] if resCalc else np.hstack([biInput, np.tile(MORinpNN, reps=[int(np.sum(fd.biDof)), 1])])
        else:
            InpuTot, biInpuTot = Input, biInput
        if not resCalc and gcoef is not None:
            d_now = diff if diff is not None else diff0
            v_now = vel if vel is not None else vel0
            tData.genSpec = self._gen_spec(Input, d_now, v_now, src)
        tData.updateData(InpuTot, biInpuTot, biLabel, gcoef, src, diff, vel, MORinpNN)
        if fresh and not resCalc:
            tData.trainDicts(fd, self.tfData)
        return tData

    # ------------------------------------------------------------------ losses and weights
    def splitLoss(self, tData, fixData=None, MORdiscArg=None, W=None):
        if not (W is None or type(W) == np.ndarray):
            raise ValueError('\'W\' must be an array with shape (,3)!')
        W = np.eye(3) if W is None else W
        fixData = self.fixData if fixData is None else fixData
        MORdiscArg = fixData.MORdiscArg if MORdiscArg is None else MORdiscArg
        comp = np.zeros([3, 1])
        lossVec = [] if fixData.lossVecflag else None
        for batch in range(fixData.MORbatchNum):
            tData = self.trainData(batch, MORdiscArg, tData)
            bc, ic, var, lv = tData.splitLoss(self.tfData, fixData.lossVecflag)
            comp += np.array([[bc, ic, var]], dtype=float).T
            if fixData.lossVecflag:
                lossVec.append(lv)
        return np.matmul(W, comp), tData, lossVec

    def trainWeight(self, weight, tData, MORdiscArg, normalizeW, useOriginalW, lossTot=1.e6):
        """Scale the weights so that the initial weighted loss equals lossTot (VarNet.py:1094-1146)."""
        td = self.PDE.timeDependent
        lossVal, tData, _ = self.splitLoss(tData, MORdiscArg=MORdiscArg)
        lossVal = np.reshape(lossVal, 3)
        terms = lossVal if td else np.array([lossVal[0], lossVal[2]])
        if useOriginalW:
            trainW = weight
        elif normalizeW:
            nw = len(weight)
            Wm = np.tile(weight, [nw, 1]) / np.reshape(weight, [nw, 1])
            trainW = np.reshape(lossTot / (np.sum(Wm, axis=1, keepdims=True) * np.reshape(terms, [nw, 1])), nw)
        else:
            trainW = lossTot / np.sum(np.array(weight) * np.array(terms)) * np.array(weight)
        if not td:
            trainW = np.array([trainW[0], 0., trainW[1]])
        trainW = np.array(trainW, dtype=float)
        msg = ('Training weight information:\n\tBC loss %.4g, IC loss %.4g, integral loss %.4g\n'
               '\trequested weights %s -> training weights %s\n' % (lossVal[0], lossVal[1], lossVal[2], weight, trainW))
        if self.trainRes.verbose:
            print(msg)
        self.trainRes.writeCase(msg)
        return trainW, tData, lossVal

    # ------------------------------------------------------------------ training
    def train(self, folderpath, weight=None, smpScheme='uniform', epochNum=500000, tol=1.e-1, verbose=True,
              saveFreq=100, pltReplace=True, saveMORdata=False, frac=None, addTrainPts=True, suppFactor=1.0,
              multiTrainUpd=False, trainUpdelay=2e4, tolUpd=0.01, reinitrain=True, updateWeights=False,
              normalizeW=False, adjustWeight=False, useOriginalW=False, batchNum=None, batchLen=None,
              shuffleData=False, shuffleFreq=1, stepsPerCall=64):
        """The reference's training loop (VarNet.py:1197-1421).  `stepsPerCall` (extension): when an epoch is one optimizer step on
        unchanged feeds (no MOR batches, one mini-batch, no shuffling, no pending re-sampling) up to that many epochs are
        taken per backend call, never across a `saveFreq` boundary; the per-epoch bookkeeping, the `tol` test and the loss
        history are applied to each returned loss in order.  If the tolerance is met inside a chunk the history stops at that
        epoch while the weights have seen the rest of the chunk (at most stepsPerCall-1 further steps).  stepsPerCall=1
        restores the reference's one-host-round-trip-per-epoch behaviour exactly."""
        if folderpath is None or is_empty(folderpath):
            raise ValueError('a folder path must be provided to backup the trained model!')
        self.folderpath = folderpath
        td = self.PDE.timeDependent
        if weight is None:
            weight = [1., 1., 1.] if td else [1., 1.]
        elif len(weight) != (3 if td else 2):
            raise ValueError('weight dimension does not match!')
        if smpScheme not in ('uniform', 'random', 'optimal'):
            raise ValueError('sampling scheme is not valid!')
        if updateWeights:
            raise NotImplementedError('periodic weight re-balancing is out of scope (latent bug in the reference, App. C.7)')
        if batchNum is None and batchLen is None and shuffleData:
            warnings.warn('shuffling data is possible for batch-optimization, setting \'shuffleData\' to False!')
            shuffleData = False
        self.smpScheme = smpScheme
        if frac is None:
            frac = 0.50 if addTrainPts else 0.25
        if not addTrainPts and np.abs(suppFactor - 1.0) > 1.e-15:
            warnings.warn('\'suppFactor\' is set to 1.0 since the number of training points does not change!')
            suppFactor = 1.0
        argDict = dict(locals())
        self.fixData.setFEdata()
        fixData, tf = self.fixData, self.tfData
        MORvar = self.PDE.MORvar
        Input, _, biInput, _ = self.trainingPoints()
        if MORvar is None:
            MORdiscArg, saveMORdata = None, False
        else:
            MORdiscArg = MORvar.discretizeArg(self.MORdiscScheme)
        tData = ManageTrainData(Input, biInput, batchNum, batchLen, saveMORdata, fixData.MORbatchNum)
        self.trainRes = TrainLog(folderpath, verbose, saveFreq)
        self.trainRes.initializeCase(self, argDict)
        trainW, tData, lossVal = self.trainWeight(weight, tData, MORdiscArg, normalizeW, useOriginalW)
        self.trainRes.trainWeight = trainW
        tData.updateDictFields('trainW', trainW)
        self.trainRes.lossComp.append(lossVal)
        tData0 = tData
        fixData0 = None
        if smpScheme == 'optimal' and addTrainPts:                      # frozen uniform tables for a comparable loss (VarNet.py:1330-1334)
            fixData0 = FIXData(self, fixData.integPnum)
            fixData0.setFEdata()
            fixData0.removeInputData()
        min_loss, epoch_time = float('inf'), 0.0
        tp_epoch, tp_updates = 1, 0
        resVal = err = lossComp = lossVec = None
        best = os.path.join(folderpath, 'best_model')
        epoch, stop = 0, False
        while epoch < epochNum and not stop:
            # epochs that can be taken in one backend call
            K = 1
            if (stepsPerCall and stepsPerCall > 1 and fixData.MORbatchNum == 1 and not shuffleData and
                    getattr(tData, 'batchNum', None) == 1 and hasattr(tData, 'optimFeedicts') and not tData.inputUpdated and
                    (smpScheme == 'uniform' or (tp_updates > 0 and not multiTrainUpd))):
                K = int(max(1, min(stepsPerCall, epochNum - epoch, saveFreq - epoch % saveFreq)))
            t0 = time.perf_counter()
            if K > 1:
                losses = tData.optimIterMany(tf, K)
            else:
                current_loss = 0
                for batch in range(fixData.MORbatchNum):
                    tData = self.trainData(batch, MORdiscArg, tData)
                    current_loss += tData.optimIter(tf)
                if hasattr(current_loss, "value"):                      # backend.Deferred: the epoch's steps are all enqueued
                    current_loss = current_loss.value()
                losses = [current_loss]
            epoch_time += time.perf_counter() - t0
            for current_loss in losses:
                epoch += 1
                if shuffleData and epoch % shuffleFreq == 0:
                    tData.shuffleTrainData(fixData)
                if epoch % saveFreq == 0:
                    if min_loss > current_loss:
                        min_loss = current_loss
                        if tf.rank == 0:
                            tf.saver.save(tf.sess, best, global_step=epoch)
                    try:
                        resVal, _, err, _ = self.residual()
                    except Exception as ex:                                 # monitoring only
                        resVal, err = None, None
                        self.trainRes.writeCase('residual monitoring unavailable: %s' % ex)
                    lossComp, _, lossVec = self.splitLoss(tData0, fixData0)
                self.trainRes.iterOutput(epoch, current_loss, min_loss, epoch_time, resVal, err, lossComp, lossVec)
                if current_loss < tol:
                    self.trainRes.writeCase('Training completed!')
                    if verbose:
                        print('Training completed!')
                    stop = True
                    break
                if smpScheme != 'uniform' and (multiTrainUpd or tp_updates == 0) and (epoch - tp_epoch) >= (trainUpdelay - 1):
                    recent = np.array(self.trainRes.loss[-5:])
                    drop = recent[:-1] - recent[1:]
                    if np.sum(drop[drop > 0]) / recent[-1] < tolUpd:          # plateau: re-sample (VarNet.py:1385-1421)
                        min_loss = float('inf')
                        tp_epoch, tp_updates = epoch, tp_updates + 1
                        self.trainRes.inpIter.append(epoch)
                        Input, _, biInput, _ = self.trainingPoints(smpScheme, frac, addTrainPts, suppFactor)
                        if addTrainPts:
                            fixData = self.fixData
                        tData = ManageTrainData(Input, biInput, batchNum, batchLen, saveMORdata, fixData.MORbatchNum)
                        self.trainRes.writeCase('Training points updated at epoch %d.' % epoch)
                        if reinitrain:
                            tf.sess.run(GlobalInit())
                            self.trainRes.writeCase('trainable variables reinitialized.')
                        if adjustWeight:
                            weight = [5 * w for w in weight[:-1]] + [weight[-1]]
                        trainW, tData, _ = self.trainWeight(weight, tData, MORdiscArg, normalizeW, useOriginalW)
                        tData.updateDictFields('trainW', trainW)
        return self.trainRes

    # ------------------------------------------------------------------ evaluation
    def _eval_input(self, x, t):
        fd, dim, td = self.fixData, self.dim, self.PDE.timeDependent
        if x is None:
            if (td and t is None) or not td:
                return fd.uniform_input
            x = fd.uniform_input[:fd.dof, :dim]
        elif np.shape(x)[1] != dim:
            raise ValueError('spatial coordinates dimension does not match domain!')
        n = np.shape(x)[0]
        if td and t is not None:
            if not (np.size(t) == 1 or np.shape(t)[0] == n):
                raise ValueError('temporal discretrization does not match spatial discretization!')
            if np.size(t) == 1:
                t = t * np.ones([n, 1])
            return np.concatenate([x, t], axis=1)
        return x

    def evaluate(self, x=None, t=None, batch=None, MORarg=None):
        """NN approximation of the solution (VarNet.py:1510-1595)."""
        fd, tf, MORvar = self.fixData, self.tfData, self.PDE.MORvar
        Input = self._eval_input(x, t)
        n = np.shape(Input)[0]
        if MORvar is not None:
            if batch is None and MORarg is None:
                raise ValueError('batch number or argument values must be given for MOR!')
            if batch is None:
                if np.shape(MORarg)[1] != tf.inpDim - fd.feDim:
                    raise ValueError('MOR argument dimension does not match the NN input size!')
                if np.shape(MORarg)[0] not in (1, n):
                    raise ValueError('MOR argument number does not match \'x\' dimension!')
                if np.shape(MORarg)[0] == 1:
                    MORarg = np.tile(MORarg, [n, 1])
            elif is_number(batch) and batch > fd.MORbatchNum - 1:
                raise ValueError('requested batch number is higher than total available batches!')
        tData = ManageTrainData(Input, biInput=[])
        if MORvar is not None and batch is None:
            tData.updateData(InpuTot=np.hstack([Input, MORarg]))
        else:
            tData = self.trainData(batch, fd.MORdiscArg, tData, resCalc=True)
        return tData.runSession(['model'], tf)[0]

    def residual(self, Input=None, tDiscIND=None, batch=None):
        """Strong-form PDE residual and solution error on a grid (VarNet.py:1599-1692):
        res = sqrt(sum(resVec^2) * prod(hVec)) averaged over MOR batches."""
        fd, PDE, dim = self.fixData, self.PDE, self.dim
        td = PDE.timeDependent
        nB = fd.MORbatchNum
        if is_number(batch) and batch is not None:
            if batch > nB - 1:
                raise ValueError('requested batch number is higher than total available batches!')
            batches, nB = range(batch, batch + 1), 1
        else:
            batches = range(nB)
        if Input is None:
            Input, cEx, diff_dx = fd.uniform_input, fd.cEx, fd.d_diff
        else:
            targs = [Input[:, dim:dim + 1]] if td else []
            cEx = PDE.cEx(Input[:, :dim], *targs) if PDE.cEx is not None else None
            diff_dx = PDE.d_diffFun(Input[:, :dim], *targs)
        res, err = 0, (0 if PDE.cEx is not None else None)
        tData = ManageTrainData(Input, biInput=None)
        resVec = cApp = None
        for b in batches:
            tData = self.trainData(b, fd.MORdiscArg, tData, resCalc=True)
            cApp, resVec = tData.runSession(['model', 'residual'], self.tfData, diff_dx=diff_dx)
            if PDE.cEx is not None:
                err += l2_err(cEx, cApp)
            res += np.sqrt(np.sum(np.asarray(resVec, dtype=float) ** 2) * np.prod(fd.hVec))
        res = res / nB
        if err is not None:
            err = err / nB
        return res, resVec, err, cApp

    # ------------------------------------------------------------------ checkpoints
    def loadModel(self, iterNum=None, folderpath=None, oldpath=None):
        """Restore the best (or a given) checkpoint written by `train` (flat-vector .npz)."""
        folderpath = getattr(self, 'folderpath', None) if folderpath is None else folderpath
        if folderpath is None:
            raise ValueError('a folder path must be provided to load the trained model!')
        if iterNum is None:
            files = glob.glob(os.path.join(folderpath, 'best_model-*.npz'))
            if not files:
                raise ValueError('no checkpoint found in %s' % folderpath)
            fname = max(files, key=lambda f: int(os.path.basename(f)[len('best_model-'):-4]))
        else:
            fname = os.path.join(folderpath, 'best_model-%d.npz' % iterNum)
        self.tfData.saver.restore(self.tfData.sess, fname)
        return fname

    def saveNNparam(self, folderpath=None):
        """Export the weights per layer (the reference writes .mat/.m, VarNet.py:2179-2260)."""
        folderpath = self.folderpath if folderpath is None else folderpath
        out = {name.replace('/', '_'): arr for name, arr in self.tfData.trainable_variables()}
        path = os.path.join(folderpath, 'NNparam.npz')
        np.savez(path, **out)
        return path
