"""The C-ABI shared library builds, loads, and exports every symbol include/ml4ca_b200.h declares.
No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ml4ca_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ml4ca_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "ml4ca_env_step" in syms and "ml4ca_env_reset" in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol(lib_built):
    handle = ctypes.CDLL(lib_built)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, "declared in include/ml4ca_b200.h but not exported: %s" % missing


def test_python_binding_covers_header(lib_built):
    from ml4ca_b200 import _lib
    assert set(_lib.exported_symbols()) == set(declared_symbols())
    assert _lib.lib().ml4ca_version().decode().endswith("sm_100a")


def test_cfg_defaults_follow_reference(lib_built):
    """customEnv.py:26,79-83,337,361,386 -- host-only entry, works without a GPU."""
    import math
    from ml4ca_b200 import _lib
    L = _lib.lib()
    cfg = _lib.EnvCfg()
    assert L.ml4ca_env_cfg_default(3, 1, 1, ctypes.byref(cfg)) == 0
    assert cfg.n_substeps == 20 and cfg.max_ep_len == 400 and abs(cfg.step_dt - 0.2) < 1e-7
    assert abs(cfg.ss_bounds[2] - math.pi / 4) < 1e-6 and abs(cfg.ss_bounds[3] - 1.4) < 1e-6
    a, o = ctypes.c_int32(), ctypes.c_int32()
    assert L.ml4ca_env_dims(ctypes.byref(cfg), ctypes.byref(a), ctypes.byref(o)) == 0 and (a.value, o.value) == (7, 9)
    for kind, cont, ext, want in [(0, 0, 1, (6, 9)), (1, 0, 0, (3, 6)), (2, 0, 1, (5, 9)), (3, 0, 0, (5, 6))]:
        assert L.ml4ca_env_cfg_default(kind, cont, ext, ctypes.byref(cfg)) == 0
        assert L.ml4ca_env_dims(ctypes.byref(cfg), ctypes.byref(a), ctypes.byref(o)) == 0
        assert (a.value, o.value) == want
    # bad arguments: error code + message, never an exception across the ABI
    assert L.ml4ca_env_cfg_default(1, 1, 0, ctypes.byref(cfg)) == -1      # cont_ang on a non-final env
    assert b"final" in L.ml4ca_last_error()
    assert L.ml4ca_env_cfg_default(1, 0, 1, ctypes.byref(cfg)) == -1      # RevoltSimple + extended state
    assert L.ml4ca_env_cfg_default(9, 0, 0, ctypes.byref(cfg)) == -1


def test_no_cpu_fallback_without_device(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from ml4ca_b200 import _lib
    L = _lib.lib()
    cfg = _lib.EnvCfg()
    assert L.ml4ca_env_cfg_default(3, 1, 1, ctypes.byref(cfg)) == 0
    h = ctypes.c_void_p()
    assert L.ml4ca_env_create(ctypes.byref(cfg), 16, 0, ctypes.byref(h)) == -3   # ML4CA_ERR_NO_DEVICE
    assert b"no CPU fallback" in L.ml4ca_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under ml4ca_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ml4ca_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "from oracle" not in text and "import oracle" not in text, f


def test_every_entry_point_is_placed_at_a_reference_seam():
    """INTEGRATION.md names every exported function next to the reference call site it replaces."""
    import re
    hdr = open(os.path.join(ROOT, "include", "ml4ca_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    funcs = sorted(set(re.findall(r"\b(ml4ca_[a-z0-9_]+)\s*\(", hdr)))
    assert len(funcs) >= 40
    missing = [f for f in funcs if f not in doc]
    assert not missing, missing
