"""ctypes binding of libeotpatch.so (C ABI declared in include/eotpatch.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EOTPATCH_LIB") or os.path.join(_HERE, "libeotpatch.so")   # override: A/B of two builds

EOT_FLAG_MASK_OUTPUT = 1
SCORE_MAX_LEVELS = 8

# every symbol include/eotpatch.h declares (tests check the export list against the header)
SYMBOLS = ["eot_last_error", "eot_version", "eot_launch_count", "eot_workspace_bytes", "eot_box_geometry", "eot_apply_fwd",
           "eot_apply_bwd", "eot_draw_transforms", "eot_brightness_match", "eot_check_workspace", "score_workspace_bytes", "score_candidate_offset", "score_max_fwd", "score_max_bwd",
           "person_nms_workspace_bytes", "person_nms", "eot_letterbox_normalize", "eot_channel_sums",
           "eot_augment_batch", "adv_u8_box_geometry", "adv_u8_print_patch", "adv_u8_workspace_bytes", "adv_u8_add_patches",
           "patch_tv_grad", "adam_clip_update", "attack_pack_scalars", "attack_step_metrics", "nhwc_bias_act_fwd", "nhwc_bias_silu_bwd",
           "nhwc_channel_scale", "nhwc_channel_dot", "nhwc_fuse_silu_fwd", "nhwc_fuse_silu_bwd"]


class EotShape(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("patch_size", ctypes.c_int32), ("num_patches", ctypes.c_int32), ("total_boxes", ctypes.c_int32),
                ("flags", ctypes.c_uint32), ("tolerance", ctypes.c_float), ("noise_amp", ctypes.c_float),
                ("min_patch_area", ctypes.c_float), ("max_scale", ctypes.c_float),
                ("patch_stride_n", ctypes.c_int64), ("patch_stride_y", ctypes.c_int64),
                ("patch_stride_x", ctypes.c_int64)]


class EotDrawConfig(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_int64), ("step", ctypes.c_int64), ("first_image", ctypes.c_int64),
                ("max_angle", ctypes.c_float), ("max_delta", ctypes.c_float), ("perspective", ctypes.c_float),
                ("scale_lo", ctypes.c_float), ("scale_span", ctypes.c_float), ("rsv", ctypes.c_float)]


class ScoreShape(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("num_levels", ctypes.c_int32), ("num_classes", ctypes.c_int32),
                ("anchors_per_loc", ctypes.c_int32), ("level_locs", ctypes.c_int32 * SCORE_MAX_LEVELS),
                ("total_anchors", ctypes.c_int32), ("image_height", ctypes.c_float),
                ("image_width", ctypes.c_float), ("min_area", ctypes.c_float)]


class NmsShape(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("total_anchors", ctypes.c_int32), ("num_levels", ctypes.c_int32),
                ("max_output_size", ctypes.c_int32), ("max_candidates", ctypes.c_int32),
                ("level_anchors", ctypes.c_int32 * SCORE_MAX_LEVELS), ("iou_threshold", ctypes.c_float),
                ("score_threshold", ctypes.c_float), ("soft_nms_sigma", ctypes.c_float), ("score_floor", ctypes.c_float),
                ("image_height", ctypes.c_float), ("image_width", ctypes.c_float)]


_lock = threading.Lock()
_lib = None


def _declare(lib):
    vp, i32, i64, f32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
    lib.eot_last_error.restype = ctypes.c_char_p
    lib.eot_last_error.argtypes = []
    lib.eot_version.restype = ctypes.c_int
    lib.eot_launch_count.restype = ctypes.c_uint64
    lib.eot_launch_count.argtypes = []
    lib.eot_workspace_bytes.argtypes = [ctypes.POINTER(EotShape), ctypes.POINTER(sz)]
    lib.eot_box_geometry.argtypes = [ctypes.POINTER(EotShape), vp, vp, vp, vp, vp, vp]
    lib.eot_apply_fwd.argtypes = [ctypes.POINTER(EotShape), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.eot_apply_bwd.argtypes = [ctypes.POINTER(EotShape), vp, vp, vp, vp, sz, vp, ctypes.c_int, vp]
    lib.eot_draw_transforms.argtypes = [ctypes.POINTER(EotDrawConfig), i32, i32, vp, vp, vp, vp]
    lib.eot_brightness_match.argtypes = [vp, i64, vp, i64, vp, vp, sz, vp]
    lib.eot_check_workspace.argtypes = [ctypes.POINTER(EotShape), vp, vp]
    lib.score_workspace_bytes.argtypes = [ctypes.POINTER(ScoreShape), ctypes.POINTER(sz)]
    lib.score_candidate_offset.argtypes = [ctypes.POINTER(ScoreShape), ctypes.POINTER(sz)]
    lib.score_max_fwd.argtypes = [ctypes.POINTER(ScoreShape), ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp, vp,
                                  vp, vp, sz, vp]
    lib.score_max_bwd.argtypes = [ctypes.POINTER(ScoreShape), ctypes.POINTER(vp), vp, vp, ctypes.POINTER(vp), vp,
                                  vp, vp, sz, vp]
    lib.person_nms_workspace_bytes.argtypes = [ctypes.POINTER(NmsShape), ctypes.POINTER(sz)]
    lib.person_nms.argtypes = [ctypes.POINTER(NmsShape), vp, ctypes.POINTER(vp), vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    f64p = ctypes.POINTER(ctypes.c_double)
    lib.eot_letterbox_normalize.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(i32), ctypes.POINTER(i32), i32, i32, i32,
                                            f64p, f64p, vp, vp, vp]
    lib.eot_channel_sums.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.eot_augment_batch.argtypes = [vp, vp, i32, i32, i32, vp, vp, f32, f32, vp]
    f64 = ctypes.c_double
    lib.adv_u8_box_geometry.argtypes = [i32, i32, f64, f64p, i32, ctypes.POINTER(i32)]
    lib.adv_u8_print_patch.argtypes = [vp, vp, i64, vp]
    lib.adv_u8_workspace_bytes.argtypes = [i32, i32, ctypes.POINTER(sz)]
    lib.adv_u8_add_patches.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, f64, f64p, i32, vp, vp, sz, vp]
    lib.nhwc_bias_act_fwd.argtypes = [vp, vp, vp, i64, i32, i32, vp]
    lib.nhwc_bias_silu_bwd.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    lib.nhwc_channel_scale.argtypes = [vp, vp, vp, f32, vp, i32, i64, i32, vp]
    lib.nhwc_channel_dot.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.nhwc_fuse_silu_fwd.argtypes = [ctypes.POINTER(vp), i32, vp, vp, i64, vp]
    lib.nhwc_fuse_silu_bwd.argtypes = [ctypes.POINTER(vp), i32, vp, vp, ctypes.POINTER(vp), i64, vp]
    lib.patch_tv_grad.argtypes = [vp, i32, f32, vp, vp, vp]
    lib.attack_pack_scalars.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.attack_step_metrics.argtypes = [vp, vp, vp, f32, f32, vp, vp]
    lib.adam_clip_update.argtypes = [vp, vp, vp, vp, i64, f32, f32, f32, f32, i64, f32, f32, vp]
    for name in SYMBOLS:
        if name not in ("eot_last_error", "eot_launch_count"):
            getattr(lib, name).restype = ctypes.c_int


def load() -> ctypes.CDLL:
    """The CUDA library, loaded once.  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the EOT patch path)")
                lib = ctypes.CDLL(LIB_PATH)
                _declare(lib)
                _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().eot_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {rc}: {msg}")
