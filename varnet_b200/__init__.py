"""varnet_b200 — B200-native weak-form residual + gradient path behind VarNet's API.

Reference names (RizaXudayi/VarNet) exported for drop-in use:
    from varnet_b200 import VarNet, ADPDE, Domain1D, PolygonDomain2D, FE, MOR, TFNN
The CUDA engine is loaded lazily by `TFNN` / `Engine`; there is no CPU fallback.
"""
from .fe import FE
from .domain import Domain, Domain1D, PolygonDomain2D, Mesh
from .pde import ADPDE
from .mor import MOR
from .tables import FIXData, ManageTrainData

__all__ = ["FE", "Domain", "Domain1D", "PolygonDomain2D", "Mesh", "ADPDE", "MOR", "FIXData", "ManageTrainData",
           "VarNet", "TFNN", "Engine"]


def __getattr__(name):          # keep `import varnet_b200` light: torch/ctypes only when needed
    if name == "VarNet":
        from .trainer import VarNet
        return VarNet
    if name == "TFNN":
        from .backend import TFNN
        return TFNN
    if name == "Engine":
        from ._capi import Engine
        return Engine
    raise AttributeError(name)
