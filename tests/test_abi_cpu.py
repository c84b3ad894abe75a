"""The C-ABI shared library loads without a GPU and exports every symbol include/sivae.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

import sivae_b200
from sivae_b200 import kernels as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sivae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sivae_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built():
    assert os.path.isfile(K.LIB_PATH), "run `python __graft_entry__.py` (build()) first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(K.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sivae.h but not exported by libsivae.so"
    assert set(declared) == set(K.EXPORTED_SYMBOLS), set(declared) ^ set(K.EXPORTED_SYMBOLS)
    lib.sivae_abi_version.restype = ctypes.c_int
    assert lib.sivae_abi_version() == 1


def test_host_side_queries_work_without_gpu():
    lib = K.load_library()
    assert lib.sivae_bn_workspace_bytes(64) > 0
    assert lib.sivae_conv3_wgrad_workspace_bytes(8, 80, 96, 80, 64, 64) > 0
    assert lib.sivae_conv3_wgrad_workspace_bytes(1, 1, 1, 1, 64, 64) > 0
    assert lib.sivae_mse_workspace_bytes(8, 614400) > 0
    assert lib.sivae_conv3_to1_workspace_bytes(64) == 27 * 16 * 64 * 2
    assert lib.sivae_launch_count() == 0
