"""Functional layer over the C ABI: torch tensors in, torch tensors out, zero-copy (data_ptr + the
current CUDA stream).  PyTorch only provides device memory and streams here; all arithmetic of the
hot path runs in libeotpatch.so.  Everything raises if the input is not on a CUDA device.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import EotShape, ScoreShape


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("the EOT patch path runs on CUDA only (no CPU fallback); got a CPU tensor")


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def params_to_tensor(params: np.ndarray, device) -> torch.Tensor:
    """[N] EotBoxParams records (numpy structured, 48 B each) -> uint8 device tensor [N,48]."""
    raw = np.ascontiguousarray(params).view(np.uint8).reshape(-1, 48)
    return torch.from_numpy(raw.copy()).to(device, non_blocking=True)


@dataclass
class PatchGeometry:
    """Shapes / hyper-parameters of one Patcher or Masker call."""
    tolerance: float = 0.2          # attacker.py:465
    noise_amp: float = 0.01         # attacker.py:426
    min_patch_area: float = 4.0     # attacker.py:347
    max_scale: float = 1.0
    serial_adjoint: bool = False    # EOT_FLAG_SERIAL_ADJOINT (memory-lean backward path)


def _shape(images: torch.Tensor, patch: torch.Tensor, n_boxes: int, g: PatchGeometry, want_mask: bool) -> EotShape:
    B, H, W, C = images.shape
    if C != 3:
        raise ValueError("images must be [B,H,W,3]")
    s = EotShape()
    s.batch, s.height, s.width = B, H, W
    if patch.dim() == 3:
        P, P2, c = patch.shape
        s.num_patches = 1
        sn, sy, sx, sc = 0, patch.stride(0), patch.stride(1), patch.stride(2)
    elif patch.dim() == 4:
        n, P, P2, c = patch.shape
        if n != B:
            raise ValueError("per-image patches must be [B,P,P,3]")
        s.num_patches = B
        sn, sy, sx, sc = patch.stride()
    else:
        raise ValueError("patch must be [P,P,3] or [B,P,P,3]")
    if P != P2 or c != 3 or sc != 1:
        raise ValueError("patch must be square, 3-channel, with unit channel stride")
    s.patch_size = P
    s.total_boxes = n_boxes
    s.flags = (_lib.EOT_FLAG_MASK_OUTPUT if want_mask else 0) | (2 if g.serial_adjoint else 0)
    s.tolerance, s.noise_amp, s.min_patch_area, s.max_scale = g.tolerance, g.noise_amp, g.min_patch_area, g.max_scale
    s.patch_stride_n, s.patch_stride_y, s.patch_stride_x = sn, sy, sx
    return s


def workspace_bytes(shape: EotShape) -> int:
    n = ctypes.c_size_t(0)
    _lib.check(_lib.load().eot_workspace_bytes(ctypes.byref(shape), ctypes.byref(n)), "eot_workspace_bytes")
    return int(n.value)


def box_geometry(images_shape: Tuple[int, int, int, int], patch_size: int, boxes: torch.Tensor,
                 offsets: torch.Tensor, params: torch.Tensor, scale: torch.Tensor,
                 g: PatchGeometry = PatchGeometry()) -> torch.Tensor:
    """`Patcher.create` + area filter + int cast for all boxes -> int32 [N,8]
    (y0,x0,ps,d,pad_lo,pad_hi,valid,span).  Reference: attacker.py:392-394,418,448-488."""
    _need_cuda(boxes, offsets, params, scale)
    B, H, W, _ = images_shape
    s = EotShape()
    s.batch, s.height, s.width, s.patch_size, s.num_patches = B, H, W, patch_size, 1
    s.total_boxes = int(boxes.shape[0])
    s.tolerance, s.noise_amp, s.min_patch_area, s.max_scale = g.tolerance, g.noise_amp, g.min_patch_area, g.max_scale
    out = torch.zeros((s.total_boxes, 8), dtype=torch.int32, device=boxes.device)
    _lib.check(_lib.load().eot_box_geometry(ctypes.byref(s), _ptr(_f32c(boxes, "boxes")), _ptr(offsets), _ptr(params),
                                            _ptr(scale), _ptr(out), _stream()), "eot_box_geometry")
    return out


@dataclass
class ApplyContext:
    """What eot_apply_bwd needs from the forward call (the 'saved tensors')."""
    shape: EotShape
    workspace: torch.Tensor
    patch: torch.Tensor
    print_wb: torch.Tensor
    generation: int = 0          # forward calls this workspace has served: a later forward invalidates the context


def apply_forward(patch: torch.Tensor, scale: torch.Tensor, images: torch.Tensor, boxes: torch.Tensor,
                  offsets: torch.Tensor, params: torch.Tensor, print_wb: torch.Tensor,
                  g: PatchGeometry = PatchGeometry(), *, want_mask: bool = False,
                  out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    """`Patcher.call` / `Masker.call` on the GPU (attacker.py:490-498, attack_detection.py:478-498).

    boxes [N,4] + offsets [B+1] int32 are the CSR form of the reference's ragged boxes; params is the
    uint8 [N,48] EotBoxParams tensor; print_wb [B,6]; scale a 0-d/1-element float32 device tensor.
    Returns (out_images, out_masks|None, ApplyContext)."""
    _need_cuda(patch, scale, images, boxes, offsets, params, print_wb)
    images = _f32c(images, "images")
    print_wb = _f32c(print_wb, "print_wb")
    boxes = _f32c(boxes, "boxes")
    if patch.dtype != torch.float32:
        raise TypeError("patch must be float32")
    if offsets.dtype != torch.int32 or offsets.numel() != images.shape[0] + 1:
        raise ValueError("offsets must be int32 [B+1]")
    n_boxes = int(boxes.shape[0])
    shape = _shape(images, patch, n_boxes, g, want_mask)
    need = workspace_bytes(shape)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=images.device)
    if out is None:
        out = torch.empty_like(images)
    mask = torch.empty_like(images) if want_mask else None
    _lib.check(_lib.load().eot_apply_fwd(ctypes.byref(shape), _ptr(patch), _ptr(scale), _ptr(images), _ptr(boxes),
                                         _ptr(offsets), _ptr(params), _ptr(print_wb), _ptr(out), _ptr(mask),
                                         _ptr(workspace), ctypes.c_size_t(workspace.numel()), _stream()),
               "eot_apply_fwd")
    gen = getattr(workspace, "_eot_generation", 0) + 1
    workspace._eot_generation = gen
    if patch.dim() == 3 and not patch.is_contiguous():
        patch = patch.contiguous()                     # the backward reads the shared patch densely: keep that copy
    return out, mask, ApplyContext(shape, workspace, patch, print_wb, gen)


def apply_backward(ctx: ApplyContext, grad_images: torch.Tensor, *, grad_patch: Optional[torch.Tensor] = None,
                   accumulate: bool = False) -> torch.Tensor:
    """dL/dpatch [P,P,3] from dL/d(out_images) (attacker.py:217)."""
    _need_cuda(grad_images)
    if ctx.shape.num_patches != 1:
        raise ValueError("the backward exists for the shared adversarial patch only (Masker carries no gradient)")
    if getattr(ctx.workspace, "_eot_generation", ctx.generation) != ctx.generation:
        raise RuntimeError("this ApplyContext is stale: its workspace has served a later eot_apply_fwd call "
                           "(give the call whose gradient is needed its own workspace)")
    grad_images = _f32c(grad_images, "grad_images")
    P = ctx.shape.patch_size
    if grad_patch is None:
        grad_patch = torch.empty((P, P, 3), dtype=torch.float32, device=grad_images.device)
        accumulate = False
    patch = ctx.patch if ctx.patch.is_contiguous() else ctx.patch.contiguous()
    _lib.check(_lib.load().eot_apply_bwd(ctypes.byref(ctx.shape), _ptr(patch), _ptr(ctx.print_wb), _ptr(grad_images),
                                         _ptr(ctx.workspace), ctypes.c_size_t(ctx.workspace.numel()),
                                         _ptr(grad_patch), int(bool(accumulate)), _stream()), "eot_apply_bwd")
    return grad_patch


def draw_transforms(seed: int, step: int, first_image: int, offsets: torch.Tensor, box_capacity: int, *,
                    max_angle: float = 20.0 * math.pi / 180.0, max_delta: float = 0.3, perspective: float = 0.0,
                    scale_range: Optional[Tuple[float, float]] = None):
    """Everything `Patcher` / `Masker` draw from TF's RNG, in one launch (eot_draw_transforms): returns
    (params uint8 [box_capacity,48], print_wb float32 [B,6]).  offsets: int32 [B+1] on the device."""
    _need_cuda(offsets)
    if offsets.dtype != torch.int32:
        raise ValueError("offsets must be int32 [B+1]")
    B = offsets.numel() - 1
    cfg = _lib.EotDrawConfig()
    cfg.seed, cfg.step, cfg.first_image = int(seed), int(step), int(first_image)
    cfg.max_angle, cfg.max_delta, cfg.perspective = float(max_angle), float(max_delta), float(perspective)
    if scale_range is None:
        cfg.scale_lo, cfg.scale_span = -1.0, 0.0
    else:
        cfg.scale_lo, cfg.scale_span = float(scale_range[0]), float(scale_range[1] - scale_range[0])
    params = torch.empty((box_capacity, 48), dtype=torch.uint8, device=offsets.device)
    print_wb = torch.empty((B, 6), dtype=torch.float32, device=offsets.device)
    _lib.check(_lib.load().eot_draw_transforms(ctypes.byref(cfg), B, int(box_capacity), _ptr(offsets), _ptr(params),
                                               _ptr(print_wb), _stream()), "eot_draw_transforms")
    return params, print_wb


def brightness_match(src: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """`BrightnessMatcher()((src, tgt))` (brightness_matcher.py:43-73): src [h,w,3], tgt [H,W,3] in [-1,1]."""
    _need_cuda(src, tgt)
    src, tgt = _f32c(src, "src"), _f32c(tgt, "tgt")
    if src.shape[-1] != 3 or tgt.shape[-1] != 3:
        raise ValueError("src and tgt must be [...,3]")
    out = torch.empty_like(src)
    ws = torch.empty(2, dtype=torch.float64, device=src.device)
    _lib.check(_lib.load().eot_brightness_match(_ptr(src), src.numel() // 3, _ptr(tgt), tgt.numel() // 3, _ptr(out),
                                                _ptr(ws), ctypes.c_size_t(16), _stream()), "eot_brightness_match")
    return out


def check_workspace(ctx: ApplyContext) -> None:
    """Synchronising check that every valid box fitted the image (tests / debugging)."""
    _lib.check(_lib.load().eot_check_workspace(ctypes.byref(ctx.shape), _ptr(ctx.workspace), _stream()),
               "eot_check_workspace")


# --------------------------------------------------------------------------------------------------
# person-score objective
# --------------------------------------------------------------------------------------------------
def _score_shape(cls_levels: Sequence[torch.Tensor], num_classes: int, H: int, W: int, min_area: float) -> ScoreShape:
    s = ScoreShape()
    s.batch = cls_levels[0].shape[0]
    s.num_levels = len(cls_levels)
    s.num_classes = num_classes
    per_loc = cls_levels[0].shape[-1] // num_classes
    s.anchors_per_loc = per_loc
    tot = 0
    for i, c in enumerate(cls_levels):
        if c.dim() != 4 or c.shape[-1] != per_loc * num_classes or c.shape[0] != s.batch:
            raise ValueError("cls level tensors must be [B,h,w,anchors*classes]")
        s.level_locs[i] = c.shape[1] * c.shape[2]
        tot += c.shape[1] * c.shape[2]
    s.total_anchors = tot * per_loc
    s.image_height, s.image_width, s.min_area = float(H), float(W), float(min_area)
    return s


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


@dataclass
class ScoreContext:
    shape: ScoreShape
    workspace: torch.Tensor
    cls_levels: Sequence[torch.Tensor]
    max_scores: torch.Tensor


def score_max_forward(cls_levels: Sequence[torch.Tensor], box_levels: Sequence[torch.Tensor], anchors: torch.Tensor,
                      image_hw: Tuple[int, int], *, num_classes: int = 90, min_area: float = 100.0):
    """pre_nms (max-reduce) + person filter + valid filter + per-image max
    (attacker.py:118-141,190; tf2/postprocess.py:104-156).  Levels are NHWC [B,h,w,9*C] / [B,h,w,36]."""
    _need_cuda(anchors, *cls_levels, *box_levels)
    cls_levels = [_f32c(c, "cls level") for c in cls_levels]
    box_levels = [_f32c(b, "box level") for b in box_levels]
    shape = _score_shape(cls_levels, num_classes, image_hw[0], image_hw[1], min_area)
    if anchors.shape != (shape.total_anchors, 4):
        raise ValueError(f"anchors must be [{shape.total_anchors},4], got {tuple(anchors.shape)}")
    n = ctypes.c_size_t(0)
    lib = _lib.load()
    _lib.check(lib.score_workspace_bytes(ctypes.byref(shape), ctypes.byref(n)), "score_workspace_bytes")
    dev = anchors.device
    ws = torch.empty(int(n.value), dtype=torch.uint8, device=dev)
    B = shape.batch
    max_scores = torch.empty(B, dtype=torch.float32, device=dev)
    argmax = torch.empty(B, dtype=torch.int32, device=dev)
    ncand = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.check(lib.score_max_fwd(ctypes.byref(shape), _ptr_array(cls_levels), _ptr_array(box_levels),
                                 _ptr(_f32c(anchors, "anchors")), _ptr(max_scores), _ptr(argmax), _ptr(ncand),
                                 _ptr(ws), ctypes.c_size_t(ws.numel()), _stream()), "score_max_fwd")
    return max_scores, argmax, ncand, ScoreContext(shape, ws, cls_levels, max_scores)


def score_candidate_view(ctx: ScoreContext) -> torch.Tensor:
    """[B,A] float32 view of the candidate scores score_max_fwd left in its workspace (score, or -1 where the
    anchor is not a person / not a valid box): the reference's ragged `scores` of attacker.py:134-139 in dense
    form; the library says where they sit (`score_candidate_offset`)."""
    B, A = ctx.shape.batch, ctx.shape.total_anchors
    o = ctypes.c_size_t(0)
    _lib.check(_lib.load().score_candidate_offset(ctypes.byref(ctx.shape), ctypes.byref(o)), "score_candidate_offset")
    off = int(o.value)
    return ctx.workspace[off: off + B * A * 4].view(torch.float32).view(B, A)


def score_max_backward(ctx: ScoreContext, scale: torch.Tensor, *, want_loss: bool = True):
    """Dense zero-filled dL/dcls per level + dL/dscale (+ data loss) for
    loss = sum(M^2 + (M-scale)^2) (attacker.py:191,193)."""
    dev = ctx.max_scores.device
    dcls = [torch.empty_like(c) for c in ctx.cls_levels]
    dscale = torch.empty((), dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev) if want_loss else None
    _lib.check(_lib.load().score_max_bwd(ctypes.byref(ctx.shape), _ptr_array(ctx.cls_levels), _ptr(ctx.max_scores),
                                         _ptr(scale), _ptr_array(dcls), _ptr(dscale), _ptr(loss),
                                         _ptr(ctx.workspace), ctypes.c_size_t(ctx.workspace.numel()), _stream()),
               "score_max_bwd")
    return dcls, dscale, loss


# --------------------------------------------------------------------------------------------------
# first-pass post-processing (NMS)
# --------------------------------------------------------------------------------------------------
@dataclass
class NmsResult:
    nms_boxes: torch.Tensor       # [B,max_out,4] selection order, zero padded
    nms_scores: torch.Tensor      # [B,max_out]
    valid_len: torch.Tensor       # [B] int32
    row_splits: torch.Tensor      # [B+1] int32
    ragged_boxes: torch.Tensor    # [B*max_out,4], first row_splits[B] rows valid
    ragged_scores: torch.Tensor   # [B*max_out]


def person_nms(cand_score: torch.Tensor, box_levels: Sequence[torch.Tensor], anchors: torch.Tensor,
               image_hw: Tuple[int, int], *, max_output_size: int = 100, iou_threshold: float = 1.0,
               score_threshold: float = 0.5, soft_nms_sigma: float = 0.25, score_floor: float = 0.5,
               max_candidates: int = 0) -> NmsResult:
    """Per-image NonMaxSuppressionV5 + clip_boxes over the score kernel's person candidates -> padded and CSR
    outputs, all on the device (attacker.py:104-116,143-170; tf2/postprocess.py:159-205).  No host sync."""
    _need_cuda(cand_score, anchors, *box_levels)
    box_levels = [_f32c(b, "box level") for b in box_levels]
    cand_score = _f32c(cand_score, "cand_score")
    B, A = cand_score.shape
    s = _lib.NmsShape()
    s.batch, s.total_anchors, s.num_levels = B, A, len(box_levels)
    s.max_output_size, s.max_candidates = int(max_output_size), int(max_candidates)
    for i, bl in enumerate(box_levels):
        if bl.shape[0] != B or bl.numel() % (B * 4):
            raise ValueError("box level tensors must be [B,h,w,anchors*4]")
        s.level_anchors[i] = bl.numel() // (B * 4)
    s.iou_threshold, s.score_threshold = float(iou_threshold), float(score_threshold)
    s.soft_nms_sigma, s.score_floor = float(soft_nms_sigma), float(score_floor)
    s.image_height, s.image_width = float(image_hw[0]), float(image_hw[1])
    if anchors.shape != (A, 4):
        raise ValueError(f"anchors must be [{A},4], got {tuple(anchors.shape)}")
    lib = _lib.load()
    n = ctypes.c_size_t(0)
    _lib.check(lib.person_nms_workspace_bytes(ctypes.byref(s), ctypes.byref(n)), "person_nms_workspace_bytes")
    dev = cand_score.device
    ws = torch.empty(int(n.value), dtype=torch.uint8, device=dev)
    M = int(max_output_size)
    r = NmsResult(torch.empty((B, M, 4), dtype=torch.float32, device=dev), torch.empty((B, M), dtype=torch.float32, device=dev),
                  torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B + 1, dtype=torch.int32, device=dev),
                  torch.empty((B * M, 4), dtype=torch.float32, device=dev), torch.empty(B * M, dtype=torch.float32, device=dev))
    _lib.check(lib.person_nms(ctypes.byref(s), _ptr(cand_score), _ptr_array(box_levels), _ptr(_f32c(anchors, "anchors")),
                              _ptr(r.nms_boxes), _ptr(r.nms_scores), _ptr(r.valid_len), _ptr(r.row_splits),
                              _ptr(r.ragged_boxes), _ptr(r.ragged_scores), _ptr(ws), ctypes.c_size_t(ws.numel()),
                              _stream()), "person_nms")
    return r


# --------------------------------------------------------------------------------------------------
# input pipeline
# --------------------------------------------------------------------------------------------------
def letterbox_normalize(frames: Sequence[torch.Tensor], output_size: Tuple[int, int], mean_rgb, stddev_rgb,
                        out: Optional[torch.Tensor] = None, want_sums: bool = True):
    """`DataSequence._map_fn` for a batch (train_data_generator.py:55-75): uint8 CUDA frames [h_i,w_i,3] of any
    sizes -> ([B,H,W,3] float32, channel sums [B,3] float64 for `augment_batch`)."""
    _need_cuda(*frames)
    B = len(frames)
    if B == 0:
        raise ValueError("no frames")
    fr = []
    for f in frames:
        if f.dtype != torch.uint8 or f.dim() != 3 or f.shape[2] != 3:
            raise TypeError("frames must be uint8 [h,w,3]")
        fr.append(f if f.is_contiguous() else f.contiguous())
    H, W = int(output_size[0]), int(output_size[1])
    dev = fr[0].device
    if out is None:
        out = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    sums = torch.empty((B, 3), dtype=torch.float64, device=dev) if want_sums else None
    heights = (ctypes.c_int32 * B)(*[int(f.shape[0]) for f in fr])
    widths = (ctypes.c_int32 * B)(*[int(f.shape[1]) for f in fr])
    mean = (ctypes.c_double * 3)(*np.broadcast_to(np.asarray(mean_rgb, np.float64), (3,)).tolist())
    std = (ctypes.c_double * 3)(*np.broadcast_to(np.asarray(stddev_rgb, np.float64), (3,)).tolist())
    _lib.check(_lib.load().eot_letterbox_normalize(_ptr_array(fr), heights, widths, B, H, W, mean, std, _ptr(out),
                                                   _ptr(sums), _stream()), "eot_letterbox_normalize")
    return out, sums


def channel_sums(images: torch.Tensor) -> torch.Tensor:
    _need_cuda(images)
    images = _f32c(images, "images")
    B, H, W, _ = images.shape
    sums = torch.empty((B, 3), dtype=torch.float64, device=images.device)
    _lib.check(_lib.load().eot_channel_sums(_ptr(images), B, H, W, _ptr(sums), _stream()), "eot_channel_sums")
    return sums


def augment_batch(images: torch.Tensor, flip: Optional[torch.Tensor], contrast_factor: float, brightness_delta: float,
                  sums: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """train_data_generator.py:218-222: flips, RandomContrast, random_brightness, clip; draws are explicit."""
    _need_cuda(images, flip, sums)
    images = _f32c(images, "images")
    B, H, W, _ = images.shape
    if sums is None:
        sums = channel_sums(images)
    if flip is not None and (flip.dtype != torch.uint8 or flip.numel() != B):
        raise TypeError("flip must be uint8 [B]")
    if out is None:
        out = torch.empty_like(images)
    _lib.check(_lib.load().eot_augment_batch(_ptr(images), _ptr(out), B, H, W, _ptr(flip), _ptr(sums),
                                             ctypes.c_float(contrast_factor), ctypes.c_float(brightness_delta),
                                             _stream()), "eot_augment_batch")
    return out


# --------------------------------------------------------------------------------------------------
# victim stand-in epilogues (not a reference interface; see csrc/victim_ops.cu)
# --------------------------------------------------------------------------------------------------
def nhwc_bias_act(x: torch.Tensor, bias: torch.Tensor, act: bool, out: torch.Tensor) -> torch.Tensor:
    """out = act(x + bias[c]) for a channels_last [N,C,H,W] float32 tensor (out may be x)."""
    N, C, H, W = x.shape
    _lib.check(_lib.load().nhwc_bias_act_fwd(_ptr(x), _ptr(bias), _ptr(out), N * H * W, C, 1 if act else 0, _stream()),
               "nhwc_bias_act_fwd")
    return out


def nhwc_bias_silu_backward(x: torch.Tensor, bias: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """dL/dx of silu(x + bias[c]) given dL/dy (channels_last float32)."""
    N, C, H, W = x.shape
    if not grad.is_contiguous(memory_format=torch.channels_last):
        grad = grad.contiguous(memory_format=torch.channels_last)
    dx = torch.empty_like(x)
    _lib.check(_lib.load().nhwc_bias_silu_bwd(_ptr(x), _ptr(bias), _ptr(grad), _ptr(dx), N * H * W, C, _stream()),
               "nhwc_bias_silu_bwd")
    return dx


def nhwc_channel_scale(y: torch.Tensor, gate: torch.Tensor, shift: Optional[torch.Tensor] = None, shift_mul: float = 0.0,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = y * gate[n,c] (+ shift[n,c] * shift_mul) for channels_last [N,C,H,W]; gate / shift are [N,C] contiguous."""
    N, C, H, W = y.shape
    if out is None:
        out = torch.empty_like(y)
    _lib.check(_lib.load().nhwc_channel_scale(_ptr(y), _ptr(gate), _ptr(shift), ctypes.c_float(shift_mul), _ptr(out), N, H * W, C,
                                              _stream()), "nhwc_channel_scale")
    return out


def nhwc_fuse_silu(xs: Sequence[torch.Tensor], w: torch.Tensor) -> torch.Tensor:
    """silu(sum_i w[i] * xs[i]) for 2 or 3 tensors of identical shape and memory layout."""
    out = torch.empty_like(xs[0])
    _lib.check(_lib.load().nhwc_fuse_silu_fwd(_ptr_array(xs), len(xs), _ptr(w), _ptr(out), out.numel(), _stream()),
               "nhwc_fuse_silu_fwd")
    return out


def nhwc_fuse_silu_backward(xs: Sequence[torch.Tensor], w: torch.Tensor, dout: torch.Tensor, need: Sequence[bool]):
    dxs = [torch.empty_like(x) if nd else None for x, nd in zip(xs, need)]
    arr = (ctypes.c_void_p * len(xs))()
    for i, d in enumerate(dxs):
        arr[i] = 0 if d is None else d.data_ptr()
    _lib.check(_lib.load().nhwc_fuse_silu_bwd(_ptr_array(xs), len(xs), _ptr(w), _ptr(dout), arr, dout.numel(), _stream()),
               "nhwc_fuse_silu_bwd")
    return dxs


def nhwc_channel_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """[N,C] = sum over H,W of a * b (channels_last [N,C,H,W]); deterministic."""
    N, C, H, W = a.shape
    out = torch.empty((N, C), dtype=torch.float32, device=a.device)
    ws = torch.empty(16 * N * C, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().nhwc_channel_dot(_ptr(a), _ptr(b), _ptr(out), _ptr(ws), N, H * W, C, _stream()), "nhwc_channel_dot")
    return out


# --------------------------------------------------------------------------------------------------
# patch update
# --------------------------------------------------------------------------------------------------
def tv_grad_(patch: torch.Tensor, grad_patch: torch.Tensor, weight: float = 1e-5, want_tv: bool = True):
    """grad_patch += weight * dTV/dpatch (attacker.py:192-193); returns TV(patch) as a device scalar."""
    _need_cuda(patch, grad_patch)
    tv = torch.empty((), dtype=torch.float32, device=patch.device) if want_tv else None
    _lib.check(_lib.load().patch_tv_grad(_ptr(_f32c(patch, "patch")), patch.shape[0], ctypes.c_float(weight),
                                         _ptr(grad_patch), _ptr(tv), _stream()), "patch_tv_grad")
    return tv


def pack_scalars_(out4: torch.Tensor, max_scores: torch.Tensor, dscale: torch.Tensor, data_loss: torch.Tensor) -> None:
    """out4 <- [dL/dscale, data loss, sum M, sum M^2] (one launch; the tail of the packed all-reduce buffer)."""
    _need_cuda(out4, max_scores, dscale, data_loss)
    _lib.check(_lib.load().attack_pack_scalars(_ptr(_f32c(max_scores, "max_scores")), int(max_scores.numel()), _ptr(dscale),
                                               _ptr(data_loss), _ptr(out4), _stream()), "attack_pack_scalars")


def step_metrics(tail4: torch.Tensor, tv: torch.Tensor, scale: torch.Tensor, global_batch: int, tv_weight: float = 1e-5):
    """[loss, scale_loss, mean_max_score, std_max_score, tv_loss, scale] as one float32 [6] device tensor (one launch)."""
    _need_cuda(tail4, tv, scale)
    out = torch.empty(6, dtype=torch.float32, device=tail4.device)
    _lib.check(_lib.load().attack_step_metrics(_ptr(tail4), _ptr(tv), _ptr(scale), ctypes.c_float(float(global_batch)),
                                               ctypes.c_float(tv_weight), _ptr(out), _stream()), "attack_step_metrics")
    return out


def tv_value(patch: torch.Tensor) -> torch.Tensor:
    """tf.image.total_variation(patch) as a device scalar (validation metrics; the gradient is not touched)."""
    scratch = torch.zeros_like(patch)
    return tv_grad_(patch, scratch, 0.0, want_tv=True)


def adam_clip_(var: torch.Tensor, m: torch.Tensor, v: torch.Tensor, grad: torch.Tensor, step: int, *, lr: float,
               lo: float, hi: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-7) -> None:
    """Keras Adam update + clip constraint in one kernel (attacker_train.py:38, attacker.py:51-54,315)."""
    _need_cuda(var, m, v, grad)
    for t in (var, m, v, grad):
        if not t.is_contiguous() or t.dtype != torch.float32:
            raise ValueError("adam_clip_ needs contiguous float32 tensors")
    _lib.check(_lib.load().adam_clip_update(_ptr(var), _ptr(m), _ptr(v), _ptr(grad), var.numel(), lr, beta1, beta2,
                                            eps, step, lo, hi, _stream()), "adam_clip_update")
