"""CPU: the C-ABI library loads and exports every symbol include/gpmpc.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "gpmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rl_gp_mpc import _cabi
    assert os.path.isfile(_cabi.lib_path()), "libgpmpc.so not built (run __graft_entry__.build())"
    lib = _cabi.load_library()
    names = header_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    # the python binding declares a prototype for every header function
    assert set(names) == set(_cabi.exported_symbols())
    assert lib.gpmpc_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    from rl_gp_mpc import _cabi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _cabi.load_library()
    h = ctypes.c_void_p()
    assert lib.gpmpc_create(ctypes.byref(h), 0) == -6      # GPMPC_ERR_NO_DEVICE
    with pytest.raises(_cabi.GpmpcError):
        _cabi.Engine()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
