"""CPU tests: the CUDA shared library loads without a GPU and exports every
symbol that include/pydisort_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

from pythonic_disort_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pydisort_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_bound_functions():
    assert set(_lib.EXPORTS) == set(declared_functions())


def test_library_exports_every_declared_symbol():
    path = _lib.build()
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert _lib.bind(path).pd_abi_version() == 2


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.pd_config) == 10 * 4
    assert ctypes.sizeof(_lib.pd_state) == 12 * ctypes.sizeof(ctypes.c_void_p)


def test_workspace_query_and_argument_validation_need_no_gpu():
    lib = _lib.bind(_lib.build())
    cfg = _lib.pd_config(1024, 60, 16, 16, 32, 16, 1, 0, 1, 1)
    assert lib.pd_workspace_bytes(ctypes.byref(cfg)) > 0
    bad = _lib.pd_config(1024, 60, 15, 16, 32, 16, 1, 0, 1, 1)  # odd NQuad
    assert lib.pd_workspace_bytes(ctypes.byref(bad)) == 0


def test_no_silent_cpu_fallback():
    import torch

    import pythonic_disort_b200 as pd
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a box without CUDA")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pd.pydisort(1.0, 0.5, 4, [1, 0.5, 0.2, 0.1, 0.0], 0.5, 1.0, 0.0)
