"""Host data tables and per-step feed management for the weak-form training path.

Mirrors of the reference's `FIXData` and `ManageTrainData`
(`/root/reference/VarNetUtility.py:204-557` and `:563-1143`): same constructors, attribute
names and call protocol, so `VarNet` (varnet_b200/trainer.py, or the reference's own
`VarNet.py`) can drive the B200 backend unchanged.  What differs is *how* a feed reaches the
device: feed values are handed to `tfData.sess.run`, which uploads an array only when its
identity changed, instead of a float64->float32 cast + H2D copy per step
(`VarNetUtility.py:1044`).

Index tables (`batchInd`, `integInd`, tower/batch slicing) are integer and bit-exact with the
reference (`VarNetUtility.py:819-838`); tests/test_tables_vs_reference.py checks them.
"""
import math
import warnings

import numpy as np

from .fe import FE
from .hostutil import is_empty, is_none, pair_rows, stack_rows


class TableView:
    """Lazy gather: the rows of `base` that belong to the test functions `tf` (each owns `integNum`
    consecutive rows; tf=None means every row), optionally followed by constant columns `extra`.

    The reference materialises these gathers on the host for every mini-batch, shuffle and MOR batch
    (`InpuTot[intgInd,:]`, VarNetUtility.py:840-844; `hstack(Input, tile(MORinp))`, VarNet.py:843-851).
    A backend that advertises `supports_table_views` keeps `base` resident on the device and receives
    only the index list / the constants; any other consumer gets the materialised array via
    `np.asarray(view)`, bit-identical to the reference's copy."""

    def __init__(self, base, tf=None, integNum=1, extra=None, gen=None):
        self.base, self.tf, self.integNum = base, tf, int(integNum)
        # optional recipe (trainer.VarNet._gen_spec) from which a backend can rebuild `base` and its sibling tables on the device
        # instead of receiving them (uniform mesh, constant coefficients: SURVEY section 8 f-2)
        self.gen = gen
        self.extra = None if extra is None or np.size(extra) == 0 else np.asarray(extra, dtype=float).reshape(1, -1)
        n = len(base) if tf is None else len(tf) * self.integNum
        self.shape = (n, base.shape[1] + (0 if self.extra is None else self.extra.shape[1]))
        self.dtype = base.dtype
        self.ndim = 2

    def __len__(self):
        return self.shape[0]

    def rows(self):
        if self.tf is None:
            return slice(None)
        return (np.asarray(self.tf).reshape(-1, 1) * self.integNum + np.arange(self.integNum)).reshape(-1)

    def materialize(self):
        out = self.base[self.rows(), :]
        if self.extra is not None:
            out = np.hstack([out, np.tile(self.extra, reps=[len(out), 1])])
        return out

    def __array__(self, dtype=None, copy=None):
        out = self.materialize()
        return out if dtype is None else out.astype(dtype)

    def __getitem__(self, key):
        return self.materialize()[key]

    def take(self, tf, integNum):
        """Restrict an all-rows view to the test functions `tf`."""
        if self.tf is not None:
            raise ValueError('view is already restricted')
        return TableView(self.base, tf, integNum, self.extra, self.gen)


class FIXData:
    """Fixed tables of one VarNet instance (mesh sizes, FE tables, monitoring grid)."""

    def __init__(self, VarNet, integPnum=2):
        PDE = VarNet.PDE
        dim, domain = VarNet.dim, PDE.domain
        td = PDE.timeDependent
        if td:
            feDim, tDiscNum = dim + 1, VarNet.tDiscNum
            ht, t_coord = VarNet.timeDisc()
        else:
            feDim, tDiscNum, t_coord = dim, 1, []
        mesh = domain.getMesh(VarNet.discNum, VarNet.bDiscNum)
        he = np.reshape(mesh.he, [dim, 1])
        hVec = np.vstack([he, ht]) if td else he                       # VarNetUtility.py:284-287
        uniform_biInput, biDof = VarNet.biTrainPoints(mesh, t_coord)
        nt = mesh.dof * tDiscNum
        lossVecflag = True
        if nt > 1e6:                                                    # coarse monitoring grid (:300-303)
            lossVecflag = False
            mesh_u = domain.getMesh(discNum=100, bDiscNum=50)
            if td:
                _, t_coord = VarNet.timeDisc(tdof=100)
            uniform_biInput, _ = VarNet.biTrainPoints(mesh_u, t_coord)
        else:
            mesh_u = mesh
        uniform_input = pair_rows(mesh_u.coordinates, t_coord) if td else mesh_u.coordinates

        MORvar = PDE.MORvar
        if MORvar is not None:
            discArg = MORvar.discretizeArg(VarNet.MORdiscScheme)
            argInd = MORvar.argIndex(discArg)
            batchNum = len(argInd)
            if batchNum > 16:
                lossVecflag = False
        else:
            discArg, argInd, batchNum = None, None, 1

        self.dim, self.feDim, self.timeDependent = dim, feDim, td
        self.integPnum = integPnum
        self.dof, self.bdof = mesh.dof, mesh.bdof
        self.biDof0, self.nt0 = biDof, nt
        self.hVec = hVec
        self.biDimVal = domain.measure                                  # :293
        self.detJvec = False
        self.lossVecflag = lossVecflag
        self.uniform_input, self.uniform_biInput = uniform_input, uniform_biInput
        self.MORbatchNum, self.MORargInd, self.MORdiscArg = batchNum, argInd, discArg
        self.cEx = self.uniform_inpData = self.d_diff = None
        for name in ("integNum", "biDof", "bDofsum", "nt", "nT", "delta", "integW", "detJ", "N", "dNx", "dNt"):
            setattr(self, name, None)
        # kept for the device-side table generator (not in the reference)
        self.mesh_coord = mesh.coordinates
        self.t_coord = t_coord
        self.fe_periodic = None

    def setInputData(self, VarNet):
        """PDE data on the monitoring grid (VarNetUtility.py:364-411)."""
        PDE, dim = VarNet.PDE, self.dim
        X = self.uniform_input
        xs = X[:, :dim]
        targs = [X[:, dim:dim + 1]] if self.timeDependent else []
        self.cEx = PDE.cEx(xs, *targs) if PDE.cEx is not None else None
        self.uniform_inpData = VarNet.PDEinpData(X) if PDE.MORvar is None else [None] * 3
        self.d_diff = PDE.d_diffFun(xs, *targs)

    def setFEdata(self):
        """FE tables for the initial uniform sampling (VarNetUtility.py:415-462)."""
        biDof = self.biDof0
        self.bDofsum = np.sum(biDof[:-1]) if self.timeDependent else np.sum(biDof)
        fe = FE(self.feDim, self.integPnum)
        integNum, nT, detJ, delta, integW, N, dN = fe.basisTot(self.nt0, self.hVec)
        self.fe_periodic = fe.periodic_tables(self.hVec)
        self.integNum, self.nT, self.detJ, self.delta, self.integW = integNum, nT, detJ, delta, integW
        self.N = N
        self.dNx = dN[:, 0:self.dim]
        self.dNt = dN[:, self.dim:self.dim + 1] if self.timeDependent else np.array([[None]])
        self.biDof = biDof
        self.nt = self.nt0

    def updateOptimData(self, frac, suppFactor):
        """Tables after adding `ceil(frac*nt0)` residual-driven test functions whose supports are
        scaled by suppFactor; detJ becomes a per-test-function vector (VarNetUtility.py:466-545)."""
        if self.nt > self.nt0:
            return
        dim = self.dim
        nt1 = math.ceil(frac * self.nt0)
        scaled = np.abs(suppFactor - 1.0) > 1.e-15
        h1 = suppFactor * self.hVec if scaled else self.hVec
        _, nT1, detJ1, _, _, N1, dN1 = FE(self.feDim, self.integPnum).basisTot(nt1, h1)
        if scaled:
            self.detJ = np.vstack([detJ1 * np.ones([nt1, 1]), self.detJ * np.ones([self.nt0, 1])])
            self.detJvec = True
        self.N = np.vstack([N1, self.N])
        self.dNx = np.vstack([dN1[:, 0:dim], self.dNx])
        if self.timeDependent:
            self.dNt = np.vstack([dN1[:, dim:dim + 1], self.dNt])
        self.biDof = [b + math.ceil(frac * b) for b in self.biDof0]
        self.bDofsum = np.sum(self.biDof[:-1]) if self.timeDependent else np.sum(self.biDof)
        self.nt = self.nt0 + nt1
        self.nT = self.nT + nT1

    def removeInputData(self):
        for name in ("uniform_input", "uniform_biInput", "cEx", "uniform_inpData", "d_diff", "MORargInd", "MORdiscArg"):
            setattr(self, name, None)


class ManageTrainData:
    """Per-step training data and the list of feed dicts (one per mini-batch)."""

    def __init__(self, Input, biInput, batchNum=None, batchLen=None, saveMORdata=False, MORbatchNum=None):
        if batchNum is not None and batchLen is not None:
            raise ValueError('Only one of batch number or length properties must be provided!')
        if batchNum is not None and not (type(batchNum) == int or batchNum < 1):
            batchNum = max(1, int(np.ceil(batchNum)))
            print('\'batchNum\' must be a positive integer, using %i' % batchNum)
        if batchLen is not None and not (type(batchLen) == int or batchLen < 1):
            batchLen = max(1, int(np.ceil(batchLen)))
            print('\'batchLen\' must be a positive integer, using %i' % batchLen)
            if batchLen < 32 or batchLen > 512:
                warnings.warn('\'batchLen\' should preferably be between 32 and 512!')
        if MORbatchNum == 1:
            saveMORdata = False
        self.Input, self.biInput = Input, biInput
        self.inputUpdated = True
        self.batchNum, self.batchLen = batchNum, batchLen
        self.saveMORdata, self.MORdataSaved = saveMORdata, False
        if saveMORdata:
            self.MORbatchNum, self.MORinp, self.MORdata, self.fieldNames = MORbatchNum, [], [], []
        self.InpuTot = self.biInpuTot = self.biLabel = self.gcoef = self.sourceVal = self.diff = self.vel = None

    # ---- data updates -----------------------------------------------------------------
    _FIELDS = (("biLabel", "biLabel"), ("gcoef", "gcoef"), ("sourceVal", "source"), ("diff", "diff"), ("vel", "vel"))

    def updateData(self, InpuTot=None, biInpuTot=None, biLabel=None, gcoef=None, sourceVal=None, diff=None,
                   vel=None, inpMOR=None):
        """Store the fields that changed; None means "keep" (VarNetUtility.py:619-656)."""
        self.inputUpdated = False
        self.InpuTot, self.biInpuTot = InpuTot, biInpuTot
        given = dict(biLabel=biLabel, gcoef=gcoef, sourceVal=sourceVal, diff=diff, vel=vel)
        names = ['InpuTot', 'biInpuTot']
        for attr, field in self._FIELDS:
            if not is_none(given[attr]):
                setattr(self, attr, given[attr])
                names.append(field)
        if hasattr(self, 'optimFeedicts'):
            self.updateDictFields(names)
        if self.saveMORdata and not self.MORdataSaved:
            self.saveMORData(names, inpMOR, biLabel, gcoef, sourceVal, diff, vel)

    def saveMORData(self, fieldnames, inpMOR, biLabel, gcoef, sourceVal, diff, vel):
        if self.MORdataSaved:
            raise ValueError('all MOR data are stored!')
        if is_none(inpMOR):
            raise ValueError('MOR input values for NN must be provided!')
        given = dict(biLabel=biLabel, gcoef=gcoef, source=sourceVal, diff=diff, vel=vel)
        self.MORinp.append(inpMOR)
        self.fieldNames.append(fieldnames)
        self.__dict__.setdefault("MORgenSpec", []).append(getattr(self, "genSpec", None))
        self.MORdata.append([given[k] for k in ('biLabel', 'gcoef', 'source', 'diff', 'vel') if k in fieldnames])
        if len(self.MORinp) == self.MORbatchNum:
            self.MORdataSaved = True

    def loadMORData(self, batch):
        if not (self.saveMORdata and self.MORdataSaved):
            raise ValueError('loadMORData() can only be called when all MOR data are saved!')
        if batch < 0 or batch > self.MORbatchNum:
            raise ValueError('batch number out of range!')
        inpMOR, names, data = self.MORinp[batch], self.fieldNames[batch], list(self.MORdata[batch])
        specs = getattr(self, "MORgenSpec", None)
        self.genSpec = specs[batch] if specs else None
        if getattr(self, "_views", False):
            self.InpuTot = TableView(self.Input, None, 1, inpMOR)           # the MOR columns stay constants
        else:
            self.InpuTot = np.hstack([self.Input, np.tile(inpMOR, reps=[len(self.Input), 1])])
        self.biInpuTot = np.hstack([self.biInput, np.tile(inpMOR, reps=[len(self.biInput), 1])])
        for key, attr in (('biLabel', 'biLabel'), ('gcoef', 'gcoef'), ('source', 'sourceVal'), ('diff', 'diff'),
                          ('vel', 'vel')):
            if key in names:
                setattr(self, attr, data.pop(0))
        self.updateDictFields(names)

    def getTrainData(self):
        return (self.InpuTot, self.biInpuTot, self.biLabel, self.gcoef, self.sourceVal)

    def getAllData(self):
        return (self.Input, self.biInput, self.InpuTot, self.biInpuTot, self.biLabel, self.gcoef, self.sourceVal,
                self.diff, self.vel)

    # ---- feed dictionaries --------------------------------------------------------------
    def _slices(self):
        """Yield (batch, tower, bInd, point indices) in the reference's order: batches outer,
        towers inner, contiguous runs of `batchLen` test functions (VarNetUtility.py:829-838)."""
        n1 = 0
        views = getattr(self, "_views", False)
        cache = self.__dict__.setdefault("_tf_cache", {})
        for bi in range(self.batchNum):
            for tower in self.compTowers:
                n0, n1 = n1, min(n1 + self.batchLen, self.nt)
                bInd = self.batchInd[n0:n1]
                if views:
                    # one index-list object per slice and shuffle epoch, shared by every field's TableView
                    key = (n0, n1, getattr(self, "_tf_version", 0))
                    if key not in cache:
                        for old in [k for k in cache if k[:2] == (n0, n1)]:
                            del cache[old]
                        cache[key] = bInd.copy()
                    bInd = cache[key]
                pts = self.integInd[bInd, :].reshape(-1) if not (views and self.integNum % 4 == 0) else None
                yield bi, tower, bInd, pts

    @staticmethod
    def _local(tower):
        return getattr(tower, "local", True)

    def _take_tf(self, table, bInd):
        """Per-test-function table (vector detJ) restricted to bInd."""
        if getattr(self, "_views", False) and self.integNum % 4 == 0:
            return TableView(table, bInd, 1)
        return table[bInd, :]

    def _take(self, table, bInd, pts, gen=None):
        """Rows of a per-point table for the test functions bInd: a lazy TableView when the backend keeps
        tables resident, else the reference's host gather `table[pts, :]`."""
        if getattr(self, "_views", False) and self.integNum % 4 == 0:
            if isinstance(table, TableView):
                return table.take(bInd, self.integNum)
            return TableView(table, bInd, self.integNum, gen=gen)
        if isinstance(table, TableView):
            table = table.materialize()
        return table[pts, :]

    def trainDicts(self, fixData, tfData):
        if hasattr(self, 'optimFeedicts'):
            raise ValueError('training dictionaries are already built!')
        batchNum, batchLen = self.batchNum, self.batchLen
        if batchNum is None and batchLen is None:
            batchNum = 1
        InpuTot, biInpuTot, biLabel, gcoef, sourceVal = self.getTrainData()
        nt, nT, integNum = fixData.nt, fixData.nT, fixData.integNum
        puNum = tfData.processorNum
        self.batchInd = np.arange(nt)                                   # :819
        self.integInd = np.arange(nT).reshape([nt, integNum])           # :820
        if batchNum is None:
            batchLen = min(batchLen, nt)
            batchNum = int(np.ceil(nt / batchLen / puNum))
        else:
            batchLen = int(np.ceil(nt / batchNum / puNum))
        self.nt, self.integNum, self.puNum = nt, integNum, puNum
        self.batchNum, self.batchLen = batchNum, batchLen
        self.compTowers = tfData.compTowers
        self._views = bool(getattr(tfData, "supports_table_views", False))
        feeds = [dict() for _ in range(batchNum)]
        for bi, tw, bInd, pts in self._slices():
            if not self._local(tw):
                continue            # another rank owns this tower: do not materialise its slice
            fd = feeds[bi]
            fd[tw.Input] = self._take(InpuTot, bInd, pts)
            fd[tw.biInput] = biInpuTot
            fd[tw.biLabel] = biLabel
            fd[tw.gcoef] = self._take(gcoef, bInd, pts, gen=getattr(self, "genSpec", None))
            fd[tw.source] = self._take(sourceVal, bInd, pts)
            fd[tw.N] = self._take(fixData.N, bInd, pts)
            fd[tw.bDof] = fixData.bDofsum
            fd[tw.intShape] = [len(bInd), integNum]
            fd[tw.integW] = fixData.integW
            fd[tw.biDimVal] = fixData.biDimVal
            fd[tw.detJvec] = fixData.detJvec
            fd[tw.dNt] = self._take(fixData.dNt, bInd, pts) if fixData.timeDependent else fixData.dNt
            fd[tw.detJ] = self._take_tf(fixData.detJ, bInd) if fixData.detJvec else fixData.detJ
        self.optimFeedicts = feeds

    def updateDictFields(self, fieldnames, trainW=None, normalizeW=True):
        if not hasattr(self, 'optimFeedicts'):
            raise ValueError('first call trainDicts() to create the training dictionaries!')
        if 'trainW' in fieldnames and trainW is None:
            raise ValueError('\'trainW\' cannot be updated by None!')
        if 'trainW' in fieldnames and normalizeW:
            # boundary/initial rows are replicated in every batch and tower: the caller's array is
            # rescaled IN PLACE, as in the reference (VarNetUtility.py:900-901)
            trainW[:-1] = trainW[:-1] / self.batchNum / self.puNum
        per_point = [k for k in ('InpuTot', 'gcoef', 'source') if k in fieldnames]
        feeds = self.optimFeedicts
        if not per_point:
            for bi in range(self.batchNum):
                for tw in self.compTowers:
                    if self._local(tw):
                        self._update_shared(feeds[bi], tw, fieldnames, trainW)
            return
        for bi, tw, bInd, pts in self._slices():
            if not self._local(tw):
                continue
            fd = feeds[bi]
            self._update_shared(fd, tw, fieldnames, trainW)
            if 'InpuTot' in fieldnames:
                fd[tw.Input] = self._take(self.InpuTot, bInd, pts)
            if 'gcoef' in fieldnames:
                fd[tw.gcoef] = self._take(self.gcoef, bInd, pts, gen=getattr(self, "genSpec", None))
            if 'source' in fieldnames:
                fd[tw.source] = self._take(self.sourceVal, bInd, pts)

    def _update_shared(self, fd, tw, fieldnames, trainW):
        if 'trainW' in fieldnames:
            fd[tw.w] = trainW
        if 'biInpuTot' in fieldnames:
            fd[tw.biInput] = self.biInpuTot
        if 'biLabel' in fieldnames:
            fd[tw.biLabel] = self.biLabel

    def shuffleTrainData(self, fixData):
        """Re-draw the test-function order and the BC/IC row order (VarNetUtility.py:957-1017)."""
        if not hasattr(self, 'optimFeedicts'):
            raise ValueError('first call trainDicts() to create the training dictionaries!')
        InpuTot, biInpuTot, biLabel, gcoef, sourceVal = self.getTrainData()
        np.random.shuffle(self.batchInd)
        self._tf_version = getattr(self, "_tf_version", 0) + 1
        biInd = np.arange(len(biLabel))
        feeds = self.optimFeedicts
        for bi, tw, bInd, pts in self._slices():
            np.random.shuffle(biInd)            # drawn for every (batch, tower) to keep the RNG stream
            if not self._local(tw):
                continue
            fd = feeds[bi]
            fd[tw.biInput] = biInpuTot[biInd, :]
            fd[tw.biLabel] = biLabel[biInd, :]
            fd[tw.Input] = self._take(InpuTot, bInd, pts)
            fd[tw.gcoef] = self._take(gcoef, bInd, pts, gen=getattr(self, "genSpec", None))
            fd[tw.source] = self._take(sourceVal, bInd, pts)
            if fixData.detJvec or self._views:
                fd[tw.N] = self._take(fixData.N, bInd, pts)
                if fixData.timeDependent:
                    fd[tw.dNt] = self._take(fixData.dNt, bInd, pts)
            if fixData.detJvec:
                fd[tw.detJ] = self._take_tf(fixData.detJ, bInd)

    # ---- session drivers ------------------------------------------------------------------
    def optimIter(self, tfData):
        """One optimizer step per mini-batch; returns the summed loss (VarNetUtility.py:1021-1047)."""
        if not hasattr(self, 'optimFeedicts'):
            raise Exception('\'trainDicts\' must be called first to construct training dictionaries!')
        total = 0
        if len(self.optimFeedicts) > 1 and hasattr(tfData.sess, "run_batches") and getattr(tfData, "batch_steps", True):
            # same steps in the same order; index-list mini-batches of one resident table go to the engine in one call.
            # The losses may still be on their way (backend.Deferred): the sum is then a Deferred too, evaluated with
            # the very same additions on first use, so the caller's next trainData() overlaps the running steps
            for val in tfData.sess.run_batches(self.optimFeedicts):
                total += val
            return total
        for fd in self.optimFeedicts:
            _, val = tfData.sess.run([tfData.optMinimize, tfData.loss], feed_dict=fd)
            total += val
        return total

    def optimIterMany(self, tfData, epochs):
        """`epochs` consecutive optimIter() calls on unchanged feeds; returns the list of their summed losses.  With a single
        mini-batch per epoch the backend takes all the steps in one call (Session.run_many)."""
        if not hasattr(self, 'optimFeedicts'):
            raise Exception('\'trainDicts\' must be called first to construct training dictionaries!')
        if len(self.optimFeedicts) == 1 and hasattr(tfData.sess, "run_many"):
            return list(tfData.sess.run_many(self.optimFeedicts[0], epochs))
        return [self.optimIter(tfData) for _ in range(int(epochs))]

    def splitLoss(self, tfData, lossVecflag):
        if not hasattr(self, 'optimFeedicts'):
            raise Exception('\'trainDicts\' must be called first to construct training dictionaries!')
        tw0 = tfData.compTowers[0]
        feeds = self.optimFeedicts
        BCloss, ICloss = tfData.sess.run([tw0.BCloss, tw0.ICloss], feed_dict=feeds[0])
        varLoss, parts = 0, []
        fetch = [tfData.varLoss, tfData.lossVec if lossVecflag else []]
        for fd in feeds:
            v, lv = tfData.sess.run(fetch, feed_dict=fd)
            varLoss += v
            parts.append(lv)
        return BCloss, ICloss, varLoss, (np.vstack(parts) if lossVecflag else None)

    def runSession(self, nodList, tfData, diff_dx=None):
        """Evaluate 'model' and/or 'residual' on the stored inputs (VarNetUtility.py:1098-1142)."""
        if is_empty(nodList):
            warnings.warn('no nodes are specified for computation!')
            return []
        tw = tfData.compTowers[0]
        fetch, feed = [], {}
        if 'model' in nodList:
            fetch.append(tfData.model(tw.Input))
            feed = {tw.Input: self.InpuTot}
        if 'residual' in nodList:
            fetch.append(tw.residual)
            feed = {tw.Input: self.InpuTot, tw.diff: self.diff, tw.vel: self.vel, tw.source: self.sourceVal,
                    tw.diff_dx: diff_dx}
        return tfData.sess.run(fetch, feed_dict=feed)
