"""Oracle restatement of `Patcher` / `Masker` / `BrightnessMatcher` and their backward.

TEST INFRASTRUCTURE (see oracle/__init__.py); parity unpinned (no reference test pins
this path).  Follows /root/reference/attacker.py:344-498 (Patcher),
/root/reference/attack_detection.py:321-498 (Masker deltas) and
/root/reference/brightness_matcher.py:25-73, op by op, in float32.

Everything the reference draws from TF's RNG is an explicit input ("transform seeds"):

  per image  print_wb[b] = (w0,w1,w2, b0,b1,b2)    attacker.py:370-371
  per box    BOX_PARAMS record:
               uy, ux   unit uniforms of the centre jitter      attacker.py:473-474
               delta    random_brightness delta, U[-.3,.3)      attacker.py:427
               cos, sin of the rotation angle, U[-20deg,20deg)  attacker.py:436
               pa, pb   projective row (0 in the reference)     SURVEY.md section 0.5
               scale    per-box scale (<0: use the shared one)  attack_detection.py:453
               key0/1   Philox key of the per-texel noise       attacker.py:426
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import tfops
from .tfops import F

BOX_PARAMS = np.dtype([("uy", "f4"), ("ux", "f4"), ("delta", "f4"), ("cos", "f4"), ("sin", "f4"),
                       ("pa", "f4"), ("pb", "f4"), ("scale", "f4"),
                       ("key0", "u4"), ("key1", "u4"), ("rsv0", "u4"), ("rsv1", "u4")])

SQRT2 = F(2.0 ** 0.5)          # attacker.py:470  (2. ** .5) as a float32 constant


@dataclass
class BoxPlan:
    """Integer placement of one patch (attacker.py:418-420, 431-434)."""
    y0: int
    x0: int
    ps: int
    d: int
    pad_lo: int        # top == left
    pad_hi: int        # bottom == right
    valid: bool
    row: np.ndarray    # float row [ymin_patch, xmin_patch, ps, ps, diag] (attacker.py:488)


def centre_clamp(c: F, extent: F, limit: F) -> F:
    """Window origin for a window of `extent` centred on `c`, kept inside [0, limit]
    (attacker.py:480-486; same logic as adv_patch.py:82-90 with extent = patch size)."""
    lo = max(c - extent / F(2.0), F(0.0))
    if lo + extent > limit:
        lo = limit - extent
    return F(lo)


def create(box: Sequence[float], scale: float, uy: float, ux: float, tolerance: float,
           H: int, W: int, min_patch_area: float = 4.0) -> BoxPlan:
    """`Patcher.create` (attacker.py:448-488) + area filter (:392-394) + int cast (:418)."""
    ymin, xmin, ymax, xmax = [F(v) for v in box]
    scale, uy, ux, tol = F(scale), F(uy), F(ux), F(tolerance)
    h = ymax - ymin
    w = xmax - xmin
    longer = max(h, w)
    ps = np.floor(longer * scale)
    diag = min(SQRT2 * ps, F(W))                         # image.shape[1] for both axes (quirk)
    # tf.random.uniform((), minval=-tol*h/2, maxval=tol*h/2) = u*(max-min)+min
    lo_y = (-tol * h) / F(2.0)
    hi_y = (tol * h) / F(2.0)
    jy = uy * (hi_y - lo_y) + lo_y
    lo_x = (-tol * w) / F(2.0)
    hi_x = (tol * w) / F(2.0)
    jx = ux * (hi_x - lo_x) + lo_x
    cy = (ymin + h / F(2.0)) + jy
    cx = (xmin + w / F(2.0)) + jx
    y0 = centre_clamp(cy, diag, F(H))
    x0 = centre_clamp(cx, diag, F(W))
    row = np.array([y0, x0, ps, ps, diag], dtype=F)
    valid = bool(ps * ps > F(min_patch_area))
    Y0, X0, PS, _, D = [int(v) for v in row]              # tf.cast(..., int32) truncates
    off = (D - PS) / 2                                     # int32 true division -> float64
    lo = int(np.floor(off))
    hi = int(np.ceil(off))
    if valid and (Y0 < 0 or X0 < 0 or Y0 + D > H or X0 + D > W or D < PS):
        raise ValueError("patch window does not fit the image (TF would fail in tf.pad / scatter)")
    return BoxPlan(Y0, X0, PS, D, lo, hi, valid, row)


# ------------------------------------------------------------------------------------
# BrightnessMatcher (brightness_matcher.py:43-73) with the print adjust in front of it
# ------------------------------------------------------------------------------------
@dataclass
class MatchState:
    q_pre: np.ndarray      # w*p+b before the clip           attacker.py:372
    ys: np.ndarray         # Y of the rescaled print-adjusted patch
    mu_s: F
    mu_t: F
    y_pre: np.ndarray      # (Ys - mu_s) + mu_t before clip   brightness_matcher.py:65
    rgb_pre: np.ndarray    # yuv_to_rgb before clip           :69
    m: np.ndarray          # matched patch in [-1, 1.0079]


def image_mean_y(image: np.ndarray) -> F:
    """mean Y of the rescaled target image (brightness_matcher.py:55,59,62,64)."""
    t = (image.astype(F) + F(1.0)) * tfops.C_127_255
    k = tfops.RGB2YUV
    yt = (t[..., 0] * k[0, 0] + t[..., 1] * k[1, 0]) + t[..., 2] * k[2, 0]
    return tfops.mean_f64(yt)


def print_and_match(patch: np.ndarray, wb: np.ndarray, mu_t: F) -> MatchState:
    """random_print_adjust (attacker.py:365-372) then BrightnessMatcher.call."""
    patch = patch.astype(F, copy=False)
    w = wb[:3].astype(F)
    b = wb[3:].astype(F)
    q_pre = w * patch + b
    q = np.clip(q_pre, F(-1.0), F(1.0))
    s = (q + F(1.0)) * tfops.C_127_255
    yuv = tfops.dot3(s, tfops.RGB2YUV)
    ys = yuv[..., 0]
    mu_s = tfops.mean_f64(ys)
    y_pre = (ys - mu_s) + mu_t
    yp = np.clip(y_pre, F(0.0), F(1.0))
    yuv2 = np.stack([yp, yuv[..., 1], yuv[..., 2]], axis=-1)
    rgb_pre = tfops.dot3(yuv2, tfops.YUV2RGB)
    m = np.clip(rgb_pre, F(0.0), F(1.0)) * tfops.C_255_127 - F(1.0)
    return MatchState(q_pre, ys, mu_s, F(mu_t), y_pre, rgb_pre, m.astype(F))


def brightness_match(src: np.ndarray, tgt: np.ndarray) -> np.ndarray:
    """`BrightnessMatcher()((src, tgt))` alone (identity print adjust)."""
    return print_and_match(src, np.array([1, 1, 1, 0, 0, 0], dtype=F), image_mean_y(tgt)).m


# ------------------------------------------------------------------------------------
# add_patch_to_image (attacker.py:405-446)
# ------------------------------------------------------------------------------------
@dataclass
class BoxState:
    plan: BoxPlan
    T: np.ndarray
    u_pre: np.ndarray      # (resize + noise) + delta, before clip   attacker.py:426-427
    R: np.ndarray          # rotated padded patch [D,D,3]            :437
    sel: np.ndarray        # where-output before the outer clip      :440


def transformed_patch(m: np.ndarray, plan: BoxPlan, prm, noise_amp: float):
    """resize -> +noise -> +delta (pre-clip), attacker.py:425-427."""
    r = tfops.aa_resize(m, plan.ps, plan.ps)
    noise = tfops.uniform_noise(plan.ps * plan.ps * 3, int(prm["key0"]), int(prm["key1"]), noise_amp)
    u_pre = (r + noise.reshape(plan.ps, plan.ps, 3)) + F(prm["delta"])
    return u_pre.astype(F)


def rotate_padded(u_pre: np.ndarray, plan: BoxPlan, prm):
    """clip -> pad(-2) -> tfa.image.rotate(bilinear, fill -2), attacker.py:428-437."""
    u = np.clip(u_pre, F(-1.0), F(1.0))
    pad = np.full((plan.d, plan.d, 3), F(-2.0), dtype=F)
    pad[plan.pad_lo:plan.pad_lo + plan.ps, plan.pad_lo:plan.pad_lo + plan.ps] = u
    T = tfops.rotation_transform(prm["cos"], prm["sin"], plan.d, prm["pa"], prm["pb"])
    return tfops.projective_bilinear(pad, T, -2.0), T


@dataclass
class ImageState:
    match: MatchState
    boxes: List[BoxState] = field(default_factory=list)


def add_patches_to_image(image: np.ndarray, patch: np.ndarray, boxes: np.ndarray, params: np.ndarray,
                         wb: np.ndarray, scale: float, *, tolerance: float = 0.2, noise_amp: float = 0.01,
                         min_patch_area: float = 4.0, want_mask: bool = False):
    """`Patcher.add_patches_to_image` (attacker.py:374-403); with want_mask the Masker
    variant (attack_detection.py:354-432).  Returns (image', mask|None, ImageState)."""
    image = image.astype(F, copy=False)
    H, W = image.shape[:2]
    st = ImageState(print_and_match(patch, wb, image_mean_y(image)))
    out = image.copy()
    mask = np.zeros_like(image) if want_mask else None
    for j in range(len(boxes)):
        prm = params[j]
        sc = prm["scale"] if prm["scale"] >= 0 else scale
        plan = create(boxes[j], sc, prm["uy"], prm["ux"], tolerance, H, W, min_patch_area)
        if not plan.valid:
            if want_mask:
                # Masker loops over the UNFILTERED count (attack_detection.py:384) and would
                # index past the filtered list; the reference fails here.
                raise ValueError("Masker: a box was filtered by min_patch_area; the reference indexes out of range")
            continue
        u_pre = transformed_patch(st.match.m, plan, prm, noise_amp)
        R, T = rotate_padded(u_pre, plan, prm)
        ys, xs = slice(plan.y0, plan.y0 + plan.d), slice(plan.x0, plan.x0 + plan.d)
        bg = out[ys, xs]
        sel = np.where(R < F(-1.0), bg, R).astype(F)
        o = np.clip(sel, F(-1.0), F(1.0))
        out[ys, xs] = o
        if want_mask:
            mask[ys, xs] = image[ys, xs] - o
        st.boxes.append(BoxState(plan, T, u_pre, R, sel))
    return out, mask, st


def patcher_forward(patch: np.ndarray, images: np.ndarray, boxes: Sequence[np.ndarray],
                    params: Sequence[np.ndarray], print_wb: np.ndarray, scale: float, **kw):
    """`Patcher.call` (attacker.py:490-498): map add_patches_to_image over the batch.

    `patch` is [P,P,3] (shared) or [B,P,P,3] (Masker training: one crop per image)."""
    outs, masks, states = [], [], []
    for b in range(images.shape[0]):
        p = patch[b] if patch.ndim == 4 else patch
        o, mk, st = add_patches_to_image(images[b], p, boxes[b], params[b], print_wb[b], scale, **kw)
        outs.append(o)
        masks.append(mk)
        states.append(st)
    want_mask = kw.get("want_mask", False)
    return np.stack(outs), (np.stack(masks) if want_mask else None), states


# ------------------------------------------------------------------------------------
# Backward: tape.gradient(loss, patch) restricted to the patcher
# ------------------------------------------------------------------------------------
def add_patches_backward(G: np.ndarray, patch: np.ndarray, wb: np.ndarray, st: ImageState, dtype=F):
    """dL/dpatch contribution of one image given G = dL/d(patched image) [H,W,3].

    Also returns dL/d(input image) as left in G after the reverse sweep, without the
    brightness-target mean term (the reference discards image gradients)."""
    G = G.astype(dtype).copy()
    P = patch.shape[0]
    g_m = np.zeros((P, P, 3), dtype=dtype)
    for bs in reversed(st.boxes):
        pl = bs.plan
        ys, xs = slice(pl.y0, pl.y0 + pl.d), slice(pl.x0, pl.x0 + pl.d)
        g_o = G[ys, xs].copy()
        G[ys, xs] = 0                                               # TensorScatterUpdate grad
        g_w = g_o * ((bs.sel >= F(-1.0)) & (bs.sel <= F(1.0)))      # outer clip  attacker.py:441
        from_R = ~(bs.R < F(-1.0))
        g_R = g_w * from_R                                          # SelectV2 grad :440
        G[ys, xs] += g_w * (~from_R)                                # background share
        g_pad = tfops.projective_bilinear_grad(g_R.astype(F), bs.T).astype(dtype)
        g_u = g_pad[pl.pad_lo:pl.pad_lo + pl.ps, pl.pad_lo:pl.pad_lo + pl.ps]   # Pad grad
        g_u = g_u * ((bs.u_pre >= F(-1.0)) & (bs.u_pre <= F(1.0)))  # inner clip  :428
        g_m += tfops.aa_resize_grad(g_u, P, P, dtype=dtype)         # ScaleAndTranslateGrad
    ms = st.match
    K = tfops.RGB2YUV.astype(dtype)
    Ki = tfops.YUV2RGB.astype(dtype)
    g_rgb = g_m * dtype(tfops.C_255_127) * ((ms.rgb_pre >= 0) & (ms.rgb_pre <= 1))
    g_yuv = g_rgb @ Ki.T                                            # rgb = yuv . K'
    gY = g_yuv[..., 0] * ((ms.y_pre >= 0) & (ms.y_pre <= 1))
    gYs = gY - gY.mean(dtype=np.float64).astype(dtype)              # d(-mean(Ys))
    g_yuv_s = np.stack([gYs, g_yuv[..., 1], g_yuv[..., 2]], axis=-1)
    g_s = g_yuv_s @ K.T                                             # yuv = s . K
    g_q = g_s * dtype(tfops.C_127_255)
    w = wb[:3].astype(dtype)
    g_p = g_q * w * ((ms.q_pre >= F(-1.0)) & (ms.q_pre <= F(1.0)))  # clip + print weights
    return g_p.astype(dtype), G


def patcher_backward(G: np.ndarray, patch: np.ndarray, print_wb: np.ndarray, states, dtype=F):
    """Sum of add_patches_backward over the batch (shared patch) -> dL/dpatch [P,P,3]."""
    g = np.zeros(patch.shape, dtype=dtype)
    for b, st in enumerate(states):
        gp, _ = add_patches_backward(G[b], patch, print_wb[b], st, dtype=dtype)
        g += gp
    return g
