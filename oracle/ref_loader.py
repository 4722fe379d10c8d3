"""TEST INFRASTRUCTURE — not product code.

Loader that lets the *reference's own* NumPy half (FiniteElement / Domain / ADPDE /
MOR / VarNetUtility.FIXData+ManageTrainData / VarNet.trainingPoints+trainData) run
unmodified in this container, where tensorflow, matplotlib and IPython are absent.

It is used ONLY to (a) validate the oracle restatement and the host mirror in
``varnet_b200`` against the reference and (b) generate the golden fixtures under
``tests/golden/`` (see ``oracle/make_golden.py``).  ``/root/reference`` does not
exist on the GPU box, so nothing on the product path, in ``-m gpu`` tests, in
``smoke()`` or in ``bench.py`` may import this module.

Recipe: stub ``matplotlib*``/``tensorflow*``/``IPython`` in
``sys.modules``; give ``Domain.Path`` a NumPy even-odd ``contains_points``;
swap ``VarNet.TFNN`` for a recorder exposing the attributes the reference's
feed-dict builder touches (``VarNetUtility.py:815-856``).
"""
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("VARNET_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "VarNet.py"))


class _Anything(types.ModuleType):
    """Module stub: any attribute is another permissive stub / no-op callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _Callable(self.__name__ + "." + name)
        setattr(self, name, obj)
        return obj


class _Callable:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        return _Callable(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Callable(self._name + "." + name)

    def __iter__(self):
        return iter(())


class NumpyPath:
    """Even-odd point-in-polygon with matplotlib.path.Path's call signature
    (``Domain.py:372-373`` uses ``Path(vertices).contains_points(x)``).  Points of
    the configs used lie strictly inside/outside, so edge semantics are benign."""

    def __init__(self, vertices, *a, **k):
        self.vertices = np.asarray(vertices, dtype=float)

    def contains_points(self, pts, *a, **k):
        pts = np.asarray(pts, dtype=float)
        v = self.vertices
        n = len(v)
        inside = np.zeros(len(pts), dtype=bool)
        x, y = pts[:, 0], pts[:, 1]
        j = n - 1
        for i in range(n):
            xi, yi = v[i]
            xj, yj = v[j]
            cond = (yi > y) != (yj > y)
            with np.errstate(divide="ignore", invalid="ignore"):
                xint = (xj - xi) * (y - yi) / (yj - yi) + xi
            inside ^= cond & (x < xint)
            j = i
        return inside


_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.path", "matplotlib.animation",
    "matplotlib.patches", "matplotlib.colors", "matplotlib.cm",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
    "mpl_toolkits.axes_grid1",
    "tensorflow", "tensorflow.keras", "tensorflow.keras.models", "tensorflow.keras.layers",
    "tensorflow.python", "tensorflow.python.client", "tensorflow.python.client.device_lib",
    "IPython",
]


class RecorderTower:
    """Feed-key holder: the reference keys its feed dicts by these attributes."""

    KEYS = ["Input", "biInput", "biLabel", "gcoef", "source", "N", "bDof", "intShape",
            "integW", "biDimVal", "detJvec", "dNt", "detJ", "w", "diff", "vel", "diff_dx",
            "residual", "BCloss", "ICloss", "lossVec", "varLoss", "loss"]

    def __init__(self, idx):
        for k in self.KEYS:
            setattr(self, k, "tower%d/%s" % (idx, k))


class RecorderTFNN:
    """Stands in for ``TFModel.TFNN`` (ctor signature ``TFModel.py:85-86``)."""

    def __init__(self, dim, inpDim, layerWidth, modelId, activationFun, timeDependent,
                 RNNdata, processors, controller, lossOpt, optimizer_name, learning_rate):
        if not isinstance(processors, list):
            processors = [processors]
        self.dim, self.inpDim, self.layerWidth = dim, inpDim, layerWidth
        self.modelId, self.activationFun = modelId, activationFun
        self.timeDependent, self.lossOpt = timeDependent, lossOpt
        self.processors, self.controller = processors, controller
        self.processorNum = len(processors)
        self.optimizer_name, self.learning_rate = optimizer_name, learning_rate
        self.depth = len(layerWidth)
        self.compTowers = [RecorderTower(i) for i in range(self.processorNum)]


_loaded = None


def load_reference():
    """Import the reference modules with stubs in place; returns a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_DIR)
    sys.dont_write_bytecode = True          # the reference directory is read-only
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["matplotlib.path"].Path = NumpyPath
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import FiniteElement, Domain, ADPDE, MOR, UtilityFunc, VarNetUtility, VarNet  # noqa: E401
    Domain.Path = NumpyPath
    VarNet.TFNN = RecorderTFNN
    ns = types.SimpleNamespace(
        FE=FiniteElement.FE, Domain1D=Domain.Domain1D, PolygonDomain2D=Domain.PolygonDomain2D,
        ADPDE=ADPDE.ADPDE, MOR=MOR.MOR, UF=UtilityFunc.UF, VarNet=VarNet.VarNet,
        FIXData=VarNetUtility.FIXData, ManageTrainData=VarNetUtility.ManageTrainData,
        modules=dict(FiniteElement=FiniteElement, Domain=Domain, ADPDE=ADPDE, MOR=MOR,
                     UtilityFunc=UtilityFunc, VarNetUtility=VarNetUtility, VarNet=VarNet))
    _loaded = ns
    return ns


def reference_feed_dicts(vn, batchNum=None, mor_batch=0):
    """Run the reference's own table pipeline for a constructed reference ``VarNet``
    and return (tData, list of per-batch feed dicts keyed by plain names).

    Mirrors the calls made by ``VarNet.train`` before the epoch loop
    (``VarNet.py:1284-1321``) without touching TensorFlow."""
    ref = load_reference()
    fd = vn.fixData
    fd.setFEdata()
    Input, _, biInput, biDof = vn.trainingPoints()
    MORvar = vn.PDE.MORvar
    if MORvar is None:
        MORdiscArg = None
    else:
        MORdiscArg = MORvar.discretizeArg(vn.MORdiscScheme)
    tData = ref.ManageTrainData(Input, biInput, batchNum, None, False, fd.MORbatchNum)
    tData = vn.trainData(mor_batch, MORdiscArg, tData)
    out = []
    for fdict in tData.optimFeedicts:
        towers = {}
        for key, val in fdict.items():
            tower, name = key.split("/")
            towers.setdefault(tower, {})[name] = val
        out.append([towers[t] for t in sorted(towers)])
    return tData, out, MORdiscArg
