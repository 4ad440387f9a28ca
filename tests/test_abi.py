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


C_LAYOUT_PROBE = r"""
#include <stddef.h>
#include <stdio.h>
#include "eotpatch.h"
#define S(T) printf("sizeof %s %zu\n", #T, sizeof(T))
#define O(T, f) printf("offsetof %s.%s %zu\n", #T, #f, offsetof(T, f))
int main(void) {
  S(EotShape); O(EotShape, batch); O(EotShape, height); O(EotShape, width); O(EotShape, patch_size); O(EotShape, num_patches);
  O(EotShape, total_boxes); O(EotShape, flags); O(EotShape, tolerance); O(EotShape, noise_amp); O(EotShape, min_patch_area);
  O(EotShape, max_scale); O(EotShape, patch_stride_n); O(EotShape, patch_stride_y); O(EotShape, patch_stride_x);
  S(EotBoxParams); O(EotBoxParams, uy); O(EotBoxParams, delta); O(EotBoxParams, cos_t); O(EotBoxParams, pa); O(EotBoxParams, scale);
  O(EotBoxParams, key0); O(EotBoxParams, key1);
  S(EotBoxGeometry); O(EotBoxGeometry, span);
  S(EotDrawConfig); O(EotDrawConfig, seed); O(EotDrawConfig, step); O(EotDrawConfig, first_image); O(EotDrawConfig, max_angle);
  O(EotDrawConfig, max_delta); O(EotDrawConfig, perspective); O(EotDrawConfig, scale_lo); O(EotDrawConfig, scale_span);
  S(ScoreShape); O(ScoreShape, level_locs); O(ScoreShape, total_anchors); O(ScoreShape, image_height); O(ScoreShape, min_area);
  S(NmsShape); O(NmsShape, max_candidates); O(NmsShape, level_anchors); O(NmsShape, iou_threshold); O(NmsShape, image_width);
  return 0;
}
"""


def test_struct_layouts_match_a_c_compiler(tmp_path):
    """include/eotpatch.h compiled as C11 by gcc: sizeof / offsetof of every struct that crosses the boundary against the
    ctypes mirrors in _lib.py and the structured dtype the Python side packs EotBoxParams with."""
    import shutil
    import subprocess
    import numpy as np
    from mladversarialobjectdetection_b200 import synth
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "probe.c"
    src.write_text(C_LAYOUT_PROBE)
    exe = tmp_path / "probe"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    mirrors = {"EotShape": _lib.EotShape, "EotDrawConfig": _lib.EotDrawConfig, "ScoreShape": _lib.ScoreShape, "NmsShape": _lib.NmsShape}
    checked = 0
    for line in out.splitlines():
        kind, name, value = line.split()
        value = int(value)
        if kind == "sizeof":
            if name in mirrors:
                assert ctypes.sizeof(mirrors[name]) == value, line
            elif name == "EotBoxParams":
                assert synth.BOX_PARAMS.itemsize == value == 48, line
            elif name == "EotBoxGeometry":
                assert value == 32, line                    # ops.box_geometry reads int32 [N,8]
        else:
            struct, field = name.split(".")
            if struct in mirrors:
                assert getattr(mirrors[struct], field).offset == value, line
            elif struct == "EotBoxParams":
                np_name = {"cos_t": "cos"}.get(field, field)
                assert synth.BOX_PARAMS.fields[np_name][1] == value, line
            elif struct == "EotBoxGeometry":
                assert value == 28, line
        checked += 1
    assert checked >= 40


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


def test_patch_tiff_is_read_back_by_an_independent_codec(tmp_path):
    """patch.tiff written by patch_io (attacker.py:341: tifffile.imwrite of the float32 patch) read back with OpenCV's
    libtiff (tifffile itself is not installed here); the reference reads it with tifffile.imread
    (attack_detection.py:57), which takes any baseline float32 TIFF."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    from mladversarialobjectdetection_b200 import patch_io
    patch = np.random.default_rng(4).uniform(-1, 1, (37, 53, 3)).astype(np.float32)
    path = str(tmp_path / "patch.tiff")
    patch_io.write_tiff_f32(path, patch)
    back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert back is not None and back.dtype == np.float32 and back.shape == patch.shape
    np.testing.assert_array_equal(back[..., ::-1], patch)           # OpenCV hands colour images back as BGR
    np.testing.assert_array_equal(patch_io.read_tiff_f32(path), patch)
    patch_io.save_weights(str(tmp_path / "w"), patch, 0.4, (123.675, 116.28, 103.53), (58.395, 57.12, 57.375))
    p2, s2 = patch_io.load_weights(str(tmp_path / "w"))
    np.testing.assert_array_equal(p2, patch)
    assert abs(s2 - 0.4) < 1e-7


def test_score_candidate_offset_is_where_the_layout_puts_the_candidates():
    """Host-only entry point: keys uint64[B] and the two int32[B] counters, each padded to 256 bytes, precede the dense
    candidate scores (csrc/score_max.cu score_layout); the offset stays inside the workspace and leaves room for [B,A]."""
    import __graft_entry__
    __graft_entry__.build()
    lib = _lib.load()

    def up(x):
        return (x + 255) // 256 * 256

    for B, locs in [(1, [16, 4, 1]), (8, [4096, 1024, 256, 64, 16]), (64, [4096, 1024, 256, 64, 16]), (33, [16384, 4096, 1024, 256, 64])]:
        s = _lib.ScoreShape()
        s.batch, s.num_levels, s.num_classes, s.anchors_per_loc = B, len(locs), 90, 9
        for i, n in enumerate(locs):
            s.level_locs[i] = n
        s.total_anchors = 9 * sum(locs)
        s.image_height = s.image_width = 512.0
        s.min_area = 100.0
        off, tot = ctypes.c_size_t(0), ctypes.c_size_t(0)
        assert lib.score_candidate_offset(ctypes.byref(s), ctypes.byref(off)) == 0
        assert lib.score_workspace_bytes(ctypes.byref(s), ctypes.byref(tot)) == 0
        assert off.value == up(up(B * 8) + 2 * B * 4)
        assert off.value + B * s.total_anchors * 4 <= tot.value
    assert lib.score_candidate_offset(ctypes.byref(s), None) != 0


def test_candidate_view_slices_the_workspace_at_the_library_offset():
    """ops.score_candidate_view on a host tensor standing in for the workspace (slicing only, no launch)."""
    import __graft_entry__
    __graft_entry__.build()
    import torch
    from mladversarialobjectdetection_b200 import ops
    lib = _lib.load()
    s = _lib.ScoreShape()
    s.batch, s.num_levels, s.num_classes, s.anchors_per_loc = 3, 2, 90, 9
    s.level_locs[0], s.level_locs[1] = 16, 4
    s.total_anchors = 9 * 20
    s.image_height = s.image_width = 64.0
    s.min_area = 100.0
    tot, off = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert lib.score_workspace_bytes(ctypes.byref(s), ctypes.byref(tot)) == 0
    assert lib.score_candidate_offset(ctypes.byref(s), ctypes.byref(off)) == 0
    ws = torch.zeros(tot.value, dtype=torch.uint8)
    ws[off.value: off.value + 3 * 180 * 4].view(torch.float32).copy_(torch.arange(3 * 180, dtype=torch.float32))
    view = ops.score_candidate_view(ops.ScoreContext(s, ws, [], torch.zeros(3)))
    assert view.shape == (3, 180) and torch.equal(view.flatten(), torch.arange(3 * 180, dtype=torch.float32))
