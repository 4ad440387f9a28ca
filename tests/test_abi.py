"""The C-ABI library loads on a CPU-only box and exports every symbol include/eotpatch.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from mladversarialobjectdetection_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "eotpatch.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|uint64_t|const char\*)\s+(\w+)\s*\(", src)))


def test_header_symbols_match_binding_list():
    assert _declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.eot_version() >= 100
    assert lib.eot_last_error() is not None


def test_struct_sizes_match_header_layout():
    # EotShape: 7 x int32/uint32 + 4 x float + pad + 3 x int64 ; EotBoxParams 48 B ; ScoreShape
    assert ctypes.sizeof(_lib.EotShape) == 72
    assert ctypes.sizeof(_lib.ScoreShape) == 4 * 4 + 4 * 8 + 4 + 3 * 4


def test_workspace_bytes_and_shape_validation_without_gpu():
    lib = _lib.load()
    s = _lib.EotShape()
    s.batch, s.height, s.width, s.patch_size, s.num_patches, s.total_boxes = 2, 64, 64, 16, 1, 5
    n = ctypes.c_size_t(0)
    assert lib.eot_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 0 and n.value > 0
    s.batch = 0
    assert lib.eot_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 2
    assert b"bad shape" in lib.eot_last_error()
    with pytest.raises(RuntimeError, match="bad shape"):
        _lib.check(2, "eot_workspace_bytes")
