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


def test_adv_u8_box_geometry_host_entry_matches_reference_create_fixture():
    """adv_u8_box_geometry is pure host code (no GPU): `_create` of adv_patch.py:61-92 through the C ABI against the
    123 boxes of the reference's own function (tests/golden/adv_patch_create.npz)."""
    import numpy as np
    lib = _lib.load()
    rows = np.load(os.path.join(ROOT, "tests", "golden", "adv_patch_create.npz"))["rows"]
    for H, W, scale, ymin, xmin, ymax, xmax, ry, rx, rph, rpw in rows:
        box = (ctypes.c_double * 4)(ymin, xmin, ymax, xmax)
        out = (ctypes.c_int32 * 4)()
        assert lib.adv_u8_box_geometry(int(H), int(W), float(scale), box, 1, out) == 0
        assert list(out) == [int(ry), int(rx), int(rph), int(rpw)]


def test_new_entry_points_validate_shapes_without_gpu():
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    s = _lib.NmsShape()
    s.batch, s.total_anchors, s.num_levels, s.max_output_size = 2, 90, 1, 100
    s.level_anchors[0] = 90
    s.score_threshold = 0.5
    assert lib.person_nms_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 0 and n.value >= 2 * 90 * 28
    s.max_output_size = 1000                                       # > 128 selections per image is not supported
    assert lib.person_nms_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 2
    assert b"max_output_size" in lib.eot_last_error()
    s.max_output_size, s.level_anchors[0] = 100, 80                # levels must add up to total_anchors
    assert lib.person_nms_workspace_bytes(ctypes.byref(s), ctypes.byref(n)) == 2
    assert lib.adv_u8_workspace_bytes(640, 640, ctypes.byref(n)) == 0 and n.value >= 640 * 640 * 3
    assert lib.adv_u8_workspace_bytes(0, 640, ctypes.byref(n)) == 2
    # NULL pointers are refused before anything is launched
    assert lib.eot_channel_sums(None, 1, 4, 4, None, None) == 1
    assert lib.nhwc_bias_act_fwd(None, None, None, 4, 4, 1, None) == 1
    assert ctypes.sizeof(_lib.NmsShape) == 5 * 4 + 8 * 4 + 6 * 4
