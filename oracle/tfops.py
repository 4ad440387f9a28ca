"""Restated TensorFlow 2.8.1 / TensorFlow-Addons 0.17.0 op semantics (float32 NumPy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The sources of these ops are NOT under
/root/reference (pip wheels: requirements.txt:4,16); each function names the upstream
kernel it restates and the reference call site that reaches it.  Every arithmetic op is
one float32 rounding (no FMA contraction), evaluated in the order of the upstream C++
so that `< -1` mask decisions (attacker.py:440) are reproducible bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32

# tensorflow/python/ops/image_ops_impl.py: _rgb_to_yuv_kernel / _yuv_to_rgb_kernel
# (brightness_matcher.py:58-59,69).  Rows = input channel, cols = output channel.
RGB2YUV = np.array([[0.299, -0.14714119, 0.61497538],
                    [0.587, -0.28886916, -0.51496512],
                    [0.114, 0.43601035, -0.10001026]], dtype=F)
YUV2RGB = np.array([[1.0, 1.0, 1.0],
                    [0.0, -0.394642334, 2.03206185],
                    [1.13988303, -0.58062185, 0.0]], dtype=F)

C_127_255 = F(127.0 / 255.0)   # brightness_matcher.py:32
C_255_127 = F(255.0 / 127.0)   # brightness_matcher.py:41


def dot3(x: np.ndarray, k: np.ndarray) -> np.ndarray:
    """`tf.tensordot(x, k, axes=[[-1],[0]])` for a 3-vector channel axis.

    Restated as ((x0*k0 + x1*k1) + x2*k2) per output channel, float32, unfused.
    """
    x = x.astype(F, copy=False)
    out = np.empty(x.shape[:-1] + (3,), dtype=F)
    for c in range(3):
        out[..., c] = (x[..., 0] * k[0, c] + x[..., 1] * k[1, c]) + x[..., 2] * k[2, c]
    return out


def mean_f64(x: np.ndarray) -> F:
    """`tf.reduce_mean` of a float32 tensor.

    TF's reduction order is device dependent and unknowable here; the oracle (and the
    CUDA kernels) accumulate in float64 and round once, which is within TF's own
    float32 reduction noise (<=1e-7 relative) and order independent.
    """
    return F(np.sum(x.astype(np.float64)) / x.size)


# ----------------------------------------------------------------------------------
# ScaleAndTranslate(kernel_type='triangle', antialias=True)  -- attacker.py:425
# tensorflow/core/kernels/image/scale_and_translate_op.cc : ComputeSpansCore,
# GatherRows / GatherColumns, and ComputeGradSpansCore for the gradient op.
# ----------------------------------------------------------------------------------
def compute_spans(out_size: int, in_size: int):
    """Spans for `tf.image.resize(..., [out,out], antialias=True)` (bilinear == triangle).

    scale = float32(out)/float32(in) (image_ops_impl.py: resize_images_v2), translate 0.
    Returns (starts[int32, out], weights[float32, out x span_size], span_size).
    """
    scale = F(out_size) / F(in_size)
    inv_scale = F(1.0 / float(scale))           # const float inv_scale = 1.0 / scale;
    inv_translate = F(-inv_scale * F(0.0))      # -inv_scale * translate
    kernel_scale = max(inv_scale, F(1.0))       # antialias=True
    radius = F(1.0)                             # triangle kernel
    span_size = min(2 * int(math.ceil(float(radius * kernel_scale))) + 1, in_size)
    one_over = F(1.0) / kernel_scale
    x = np.arange(out_size, dtype=F)
    col_f = x + F(0.5)
    sample_f = col_f * inv_scale + inv_translate
    outside = (sample_f < 0) | (sample_f > F(in_size))
    rk = radius * kernel_scale
    span_start = np.ceil((sample_f - rk) - F(0.5)).astype(np.int64)
    span_end = np.floor((sample_f + rk) - F(0.5)).astype(np.int64)
    span_start = np.clip(span_start, 0, in_size - 1)
    span_end = np.clip(span_end, 0, in_size - 1) + 1
    this_span = span_end - span_start
    assert int(this_span.max()) <= span_size, "span exceeds span_size (TF: errors::Internal)"
    weights = np.zeros((out_size, span_size), dtype=F)
    total = np.zeros(out_size, dtype=F)
    raw = np.zeros((out_size, span_size), dtype=F)
    for k in range(span_size):
        src = (span_start + k).astype(F)
        kernel_pos = (src + F(0.5)) - sample_f
        a = np.abs(kernel_pos * one_over)
        w = np.where(a < F(1.0), F(1.0) - a, F(0.0)).astype(F)
        w = np.where(k < this_span, w, F(0.0)).astype(F)
        raw[:, k] = w
        total = (total + w).astype(F)           # sequential float32 sum
    ok = np.abs(total) >= F(1000.0) * np.finfo(F).tiny
    inv_total = np.where(ok, F(1.0) / np.where(ok, total, F(1.0)), F(0.0)).astype(F)
    weights = (raw * inv_total[:, None]).astype(F)
    starts = span_start.astype(np.int32)
    starts[outside] = 0
    weights[outside] = 0
    return starts, weights, span_size


def _gather_axis0(x: np.ndarray, starts, weights, span_size) -> np.ndarray:
    """GatherRows: out[o] = sum_k w[o,k] * x[start[o]+k], sequential, from 0, unfused."""
    n_in = x.shape[0]
    out = np.zeros((len(starts),) + x.shape[1:], dtype=F)
    bshape = (-1,) + (1,) * (x.ndim - 1)
    for k in range(span_size):
        idx = starts.astype(np.int64) + k
        inb = idx < n_in
        w = np.where(inb, weights[:, k], F(0.0)).astype(F)
        idx = np.minimum(idx, n_in - 1)
        out = (out + w.reshape(bshape) * x[idx]).astype(F)
    return out


def aa_resize(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`tf.image.resize(x[H,W,C], [out_h,out_w], antialias=True)` (attacker.py:425).

    Rows first into a float32 [out_h, W, C] intermediate, then columns
    (scale_and_translate_op.cc: GatherSpans functor).
    """
    x = x.astype(F, copy=False)
    rs, rw, rn = compute_spans(out_h, x.shape[0])
    cs, cw, cn = compute_spans(out_w, x.shape[1])
    inter = _gather_axis0(x, rs, rw, rn)                                  # [out_h, W, C]
    out = _gather_axis0(np.ascontiguousarray(inter.transpose(1, 0, 2)), cs, cw, cn)
    return np.ascontiguousarray(out.transpose(1, 0, 2))                   # [out_h, out_w, C]


def _scatter_axis0(g: np.ndarray, starts, weights, span_size, n_in: int, dtype) -> np.ndarray:
    """Transpose of _gather_axis0 (ScaleAndTranslateGrad: ComputeGradSpansCore)."""
    out = np.zeros((n_in,) + g.shape[1:], dtype=dtype)
    bshape = (-1,) + (1,) * (g.ndim - 1)
    for k in range(span_size):
        idx = starts.astype(np.int64) + k
        inb = idx < n_in
        w = np.where(inb, weights[:, k], 0).astype(dtype)
        np.add.at(out, np.minimum(idx, n_in - 1), w.reshape(bshape) * g.astype(dtype))
    return out


def aa_resize_grad(g: np.ndarray, in_h: int, in_w: int, dtype=F) -> np.ndarray:
    """`ScaleAndTranslateGrad`: exact transpose of aa_resize. g [out_h,out_w,C] -> [in_h,in_w,C]."""
    rs, rw, rn = compute_spans(g.shape[0], in_h)
    cs, cw, cn = compute_spans(g.shape[1], in_w)
    t = _scatter_axis0(g, rs, rw, rn, in_h, dtype)                        # [in_h, out_w, C]
    t = _scatter_axis0(np.ascontiguousarray(t.transpose(1, 0, 2)), cs, cw, cn, in_w, dtype)
    return np.ascontiguousarray(t.transpose(1, 0, 2))


# ----------------------------------------------------------------------------------
# tfa.image.rotate -> ImageProjectiveTransformV3(BILINEAR, CONSTANT)  -- attacker.py:437
# tensorflow_addons/image/transform_ops.py: angles_to_projective_transforms
# tensorflow/core/kernels/image/image_ops.h: ProjectiveGenerator / bilinear_interpolation
# ----------------------------------------------------------------------------------
def rotation_transform(cos_t: F, sin_t: F, size: int, pa: F = F(0.0), pb: F = F(0.0)) -> np.ndarray:
    """8-parameter output->input transform of `tfa.image.rotate` for a size x size image.

    (cos, sin) are float32 inputs (the caller evaluates them once per box); the two
    offsets follow angles_to_projective_transforms in float32.  (pa, pb) fill the
    projective row (0 for the reference's pure rotation).
    """
    cos_t, sin_t = F(cos_t), F(sin_t)
    wm1 = F(size - 1)
    hm1 = F(size - 1)
    x_off = (wm1 - (cos_t * wm1 - sin_t * hm1)) / F(2.0)
    y_off = (hm1 - (sin_t * wm1 + cos_t * hm1)) / F(2.0)
    return np.array([cos_t, -sin_t, x_off, sin_t, cos_t, y_off, F(pa), F(pb)], dtype=F)


def invert_transform(t: np.ndarray) -> np.ndarray:
    """flat_transforms_to_matrices -> matrix_inverse -> matrices_to_flat_transforms
    (tensorflow/python/ops/image_ops.py: _image_projective_transform_v3_grad).

    Evaluated by the adjugate in float64 and rounded to float32 once (TF inverts in
    float32 by LU; the difference is at the 1e-7 level, far inside the 1e-4 gradient bar).
    """
    a, b, c, d, e, f, g, h = [float(v) for v in t]
    i = 1.0
    A = e * i - f * h
    B = -(d * i - f * g)
    C = d * h - e * g
    D = -(b * i - c * h)
    E = a * i - c * g
    Fm = -(a * h - b * g)
    G = b * f - c * e
    H = -(a * f - c * d)
    I = a * e - b * d
    # inverse = adj / det ; adj = [[A, D, G], [B, E, H], [C, Fm, I]] ; normalise by [2,2]
    return np.array([A / I, D / I, G / I, B / I, E / I, H / I, C / I, Fm / I], dtype=F)


def projective_coords(t: np.ndarray, out_h: int, out_w: int):
    """Input-space sampling coordinates of every output pixel (float32, unfused)."""
    t = t.astype(F)
    xs = np.arange(out_w, dtype=F)[None, :]
    ys = np.arange(out_h, dtype=F)[:, None]
    proj = (t[6] * xs + t[7] * ys) + F(1.0)
    ix = ((t[0] * xs + t[1] * ys) + t[2]) / proj
    iy = ((t[3] * xs + t[4] * ys) + t[5]) / proj
    return ix.astype(F), iy.astype(F), proj.astype(F)


def projective_bilinear(img: np.ndarray, t: np.ndarray, fill: float, out_h=None, out_w=None) -> np.ndarray:
    """`ImageProjectiveTransformV3(img[H,W,C], t, BILINEAR, CONSTANT, fill)`."""
    img = img.astype(F, copy=False)
    H, W = img.shape[:2]
    out_h = H if out_h is None else out_h
    out_w = W if out_w is None else out_w
    ix, iy, proj = projective_coords(t, out_h, out_w)
    fill = F(fill)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + F(1.0)
    y1 = y0 + F(1.0)

    def read(yf, xf):
        yi = yf.astype(np.int64)
        xi = xf.astype(np.int64)
        inb = (yi >= 0) & (yi < H) & (xi >= 0) & (xi < W)
        v = img[np.clip(yi, 0, H - 1), np.clip(xi, 0, W - 1)]
        return np.where(inb[..., None], v, fill).astype(F)

    wx1 = (x1 - ix)[..., None]
    wx0 = (ix - x0)[..., None]
    wy1 = (y1 - iy)[..., None]
    wy0 = (iy - y0)[..., None]
    v0 = wx1 * read(y0, x0) + wx0 * read(y0, x1)
    v1 = wx1 * read(y1, x0) + wx0 * read(y1, x1)
    out = (wy1 * v0 + wy0 * v1).astype(F)
    out = np.where((proj == 0)[..., None], fill, out).astype(F)
    return out


def projective_bilinear_grad(g: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Registered gradient of ImageProjectiveTransformV3 w.r.t. the image: the SAME
    bilinear warp applied to the upstream gradient with the inverted transform and
    fill 0 -- NOT the scatter adjoint."""
    return projective_bilinear(g, invert_transform(t), 0.0)


# ----------------------------------------------------------------------------------
# Counter-based noise (replaces tf.random.uniform at attacker.py:426; the TF stream
# itself is not reproducible without TF, so the transform seeds are explicit inputs)
# ----------------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr0: np.ndarray, key0: int, key1: int) -> np.ndarray:
    """Philox4x32-10 of counters (ctr0, 0, 0, 0) under key (key0, key1) -> [..., 4] uint32."""
    c0 = ctr0.astype(np.uint32)
    c1 = np.zeros_like(c0)
    c2 = np.zeros_like(c0)
    c3 = np.zeros_like(c0)
    k0 = np.uint32(key0 & 0xFFFFFFFF)
    k1 = np.uint32(key1 & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for r in range(10):
            if r > 0:
                k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
                k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return np.stack([c0, c1, c2, c3], axis=-1)


def uniform_noise(n_elems: int, key0: int, key1: int, amp: float) -> np.ndarray:
    """n_elems float32 draws of U[-amp, amp): element e uses word e%4 of counter e//4.

    uint32 -> float follows TF's Uint32ToFloat (23 mantissa bits, [1,2) - 1), and the
    range map is `u * (maxval - minval) + minval` as in random_ops.random_uniform.
    """
    n_ctr = (n_elems + 3) // 4
    words = philox4x32_10(np.arange(n_ctr, dtype=np.uint32), key0, key1).reshape(-1)[:n_elems]
    bits = (words & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)
    u = bits.view(F) - F(1.0)
    lo = F(-amp)
    rng = F(amp) - lo
    return (u * rng + lo).astype(F)


def total_variation(x: np.ndarray):
    """`tf.image.total_variation(x[H,W,C])` and its gradient via sign (attacker.py:192)."""
    x = x.astype(F, copy=False)
    dy = x[1:] - x[:-1]
    dx = x[:, 1:] - x[:, :-1]
    tv = F(np.sum(np.abs(dy).astype(np.float64)) + np.sum(np.abs(dx).astype(np.float64)))
    g = np.zeros_like(x)
    sy = np.sign(dy)
    sx = np.sign(dx)
    g[1:] += sy
    g[:-1] -= sy
    g[:, 1:] += sx
    g[:, :-1] -= sx
    return tv, g
