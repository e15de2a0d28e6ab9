"""The C-ABI shared library loads and exports every symbol include/mugiq_b200.h declares (no GPU needed:
nothing here launches a kernel)."""
import ctypes
import os
import re

import numpy as np

from conftest import ROOT
from mugiq_b200 import _lib

HEADER = os.path.join(ROOT, "include", "mugiq_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mugiq_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert _lib.load().mugiq_b200_version() >= 100


def test_no_torch_types_in_abi():
    text = open(HEADER).read()
    assert "torch" not in text and "at::" not in text and "std::" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def test_gamma_tables_match_oracle(oracle):
    """Host-only entry point: the tables compiled into the kernels equal the oracle's restatement of
    include/gamma.h."""
    from mugiq_b200 import ops
    for a, b in zip(ops.gamma_tables(), oracle.gamma_tables()):
        assert np.array_equal(a, b)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    g = _lib.make_geom((3, 4, 4, 4))
    rc = lib.mugiq_b200_loop_workspace_bytes(ctypes.byref(g), 4, _lib.entry_array([]), 0)
    assert rc == -1 and b"must be even" in lib.mugiq_b200_last_error()
    g = _lib.make_geom((4, 4, 4, 4), precision=2)
    assert lib.mugiq_b200_loop_workspace_bytes(ctypes.byref(g), 4, _lib.entry_array([]), 0) == -1
    g = _lib.make_geom((4, 4, 4, 4))
    assert lib.mugiq_b200_loop_workspace_bytes(ctypes.byref(g), 4, _lib.entry_array([(4, 1, 1, 1)]), 1) == -1
    assert lib.mugiq_b200_loop_workspace_bytes(ctypes.byref(g), 4, _lib.entry_array([(0, 1, 3, 1)]), 1) == -1
    assert lib.mugiq_b200_loop_workspace_bytes(ctypes.byref(g), 4, _lib.entry_array([]), 0) == 0
    assert lib.mugiq_b200_loop_workspace_bytes(ctypes.byref(g), 4, _lib.entry_array([(0, 1, 1, 2)]), 1) > 0
    assert lib.mugiq_b200_momproj_workspace_bytes(0, 1, 1, 8) == -1
    assert lib.mugiq_b200_momproj_workspace_bytes(4608, 33, 4096, 8) > 0
