"""CPU: the C-ABI library builds, loads and exports every symbol include/varnet_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "varnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = header_functions()
    for must in ("vn_create", "vn_destroy", "vn_upload_points_f64", "vn_upload_bic_f64", "vn_loss", "vn_loss_grad",
                 "vn_optimizer_step", "vn_train_step", "vn_eval_f64", "vn_residual_f64", "vn_grad_buffer"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from varnet_b200 import build, _capi
    build.build(verbose=False)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), "missing export: " + name
    assert set(header_functions()) == set(_capi.SIGNATURES), "ctypes binding and header disagree"


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from varnet_b200._capi import Engine, EngineError
    with pytest.raises(EngineError) as ei:
        Engine(1, 2, [20])
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)


def test_bad_arguments_are_rejected_before_any_device_work():
    from varnet_b200._capi import Engine
    with pytest.raises(ValueError):
        Engine(1, 2, [20], activation="relu")
    with pytest.raises(ValueError):
        Engine(1, 2, [20], optimizer="sgd")
