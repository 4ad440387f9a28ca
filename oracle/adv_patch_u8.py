"""CPU restatement of the uint8 inference twin `adv_patch.AdversarialPatch` -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/adv_patch.py:40-201 (`print_patch`, `_create`, `rescale`, `brightness_match`, `resize`,
`get_transformed_patch`, `add_adv_to_img`).  The arithmetic of `cv2.cvtColor` and `cv2.resize` lives in
opencv-python (requirements.txt:11 pins 4.5.5.64; this image has 4.13) and is restated here from
modules/imgproc/src/color_yuv.simd.hpp (8-bit fixed point, shift 14) and resize.cpp (INTER_AREA: integer box sums
for integer ratios, float32 tap tables otherwise; INTER_LINEAR on 8-bit data: 11-bit fixed point, horizontal then
vertical).

PINNED, bit for bit:
  * against cv2 itself where it is installed: all 2^24 colours through both colour conversions, random frames
    through both resizes (tests/test_oracle_adv_patch_u8.py);
  * against the reference's own `AdversarialPatch.add_adv_to_img` run in the build container on seeded frames
    (fixtures tests/golden/adv_patch_u8.npz, generator tests/golden/make_golden.py).
INTER_CUBIC (patch up-sampling, adv_patch.py:158-160): `resize_cubic_u8` restates OpenCV's own 8-bit bicubic
(resize.cpp: interpolateCubic with A = -0.75, coefficients quantised to 11 bits, integer horizontal pass, vertical pass
in float32 for the vectorised part of a row and with the 22-bit rounding shift for its tail) and is bit-identical to cv2.resize of the installed OpenCV 4.13 with IPP switched off.
The pip wheels (the reference pins opencv-python 4.5.5.64) route this one call through Intel IPP's closed-source
ippiResizeCubic, whose output differs from OpenCV's own kernel by one grey level on about 4 % of the elements and is a
float evaluation nobody can restate bit for bit (the closest open formulation found here -- float32 separable, double
coordinates -- still misses ~1e-5 of the elements); the pinned target is therefore OpenCV's kernel, and the IPP build is
reproduced to within one grey level (tests/test_oracle_adv_patch_u8.py states both).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


# ---- cv2.cvtColor, 8-bit (color_yuv.simd.hpp: RGB2YCrCb_i / YCrCb2RGB_i with the YUV coefficient set) --------
def _descale(x, n=14):
    return (x + (1 << (n - 1))) >> n


def rgb2yuv(img: np.ndarray) -> np.ndarray:
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    y = _descale(r * 4899 + g * 9617 + b * 1868)
    v = _descale((r - y) * 14369 + (128 << 14))
    u = _descale((b - y) * 8061 + (128 << 14))
    return np.stack([np.clip(y, 0, 255), np.clip(u, 0, 255), np.clip(v, 0, 255)], -1).astype(np.uint8)


def yuv2rgb(img: np.ndarray) -> np.ndarray:
    y, u, v = (img[..., i].astype(np.int64) for i in range(3))
    b = y + _descale((u - 128) * 33292)
    g = y + _descale((u - 128) * (-6472) + (v - 128) * (-9519))
    r = y + _descale((v - 128) * 18678)
    return np.stack([np.clip(r, 0, 255), np.clip(g, 0, 255), np.clip(b, 0, 255)], -1).astype(np.uint8)


# ---- cv2.resize, 8-bit --------------------------------------------------------------------------------------
def area_tab(ssize: int, dsize: int, scale: float):
    """computeResizeAreaTab: (dst index, src index, float32 weight) in table order."""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, F((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, F(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, F(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def _round_u8(x):
    return np.clip(np.rint(x), 0, 255).astype(np.uint8)          # saturate_cast<uchar>(float): round half to even


def resize_area_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA) for a down-scale."""
    sh, sw, c = img.shape
    scale_x, scale_y = sw / dw, sh / dh
    ix, iy = int(round(scale_x)), int(round(scale_y))
    if abs(scale_x - ix) < 2.220446049250313e-16 and abs(scale_y - iy) < 2.220446049250313e-16:
        s = img.reshape(dh, iy, dw, ix, c).astype(np.int64).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        return _round_u8(s.astype(F) * F(1.0 / (ix * iy)))
    xtab, ytab = area_tab(sw, dw, scale_x), area_tab(sh, dh, scale_y)
    S = img.astype(F)
    rows = {}

    def hrow(sy):
        buf = np.zeros((dw, c), F)
        for dx, sx, a in xtab:
            buf[dx] = buf[dx] + S[sy, sx] * a
        return buf
    out = np.zeros((dh, dw, c), np.uint8)
    prev = ytab[0][0]
    acc = np.zeros((dw, c), F)
    for dy, sy, beta in ytab:
        if sy not in rows:
            rows[sy] = hrow(sy)
        if dy != prev:
            out[prev] = _round_u8(acc)
            acc = beta * rows[sy]
            prev = dy
        else:
            acc = acc + beta * rows[sy]
    out[prev] = _round_u8(acc)
    return out


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh)) (INTER_LINEAR) on 8-bit data."""
    sh, sw, c = img.shape
    if (sh, sw) == (dh, dw):
        return img.copy()
    if sw == 2 * dw and sh == 2 * dh:                               # routed to the 2x2 box average
        s = img.reshape(dh, 2, dw, 2, c).astype(np.int64).sum(axis=(1, 3))
        return ((s + 2) >> 2).astype(np.uint8)

    def taps(d, s, fix):
        scale = 1.0 / (d / s)
        f = ((np.arange(d, dtype=np.float64) + 0.5) * scale - 0.5).astype(F)
        sx = np.floor(f).astype(np.int64)
        f = (f - sx.astype(F)).astype(F)
        if fix:                                                     # horizontal taps only; rows are clamped instead
            lo = sx < 0
            f[lo] = 0
            sx[lo] = 0
            hi = sx >= s - 1
            f[hi] = 0
            sx[hi] = s - 1
        return sx, np.rint((F(1) - f) * F(2048)).astype(np.int64), np.rint(f * F(2048)).astype(np.int64)
    sx, a0, a1 = taps(dw, sw, True)
    sy, b0, b1 = taps(dh, sh, False)
    S = img.astype(np.int64)
    H = S[:, sx] * a0[None, :, None] + S[:, np.minimum(sx + 1, sw - 1)] * a1[None, :, None]
    r0, r1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    v = (((b0[:, None, None] * (H[r0] >> 4)) >> 16) + ((b1[:, None, None] * (H[r1] >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


# ---- adv_patch.AdversarialPatch -----------------------------------------------------------------------------
def _cubic_tab(dsize: int, ssize: int):
    """resize.cpp (INTER_CUBIC): source offset and the four 11-bit fixed-point taps of every destination index."""
    scale = ssize / dsize
    ofs = np.zeros(dsize, np.int64)
    taps = np.zeros((dsize, 4), np.int64)
    A = np.float32(-0.75)
    one = np.float32(1.0)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s0 = int(np.floor(f))
        x = np.float32(f - np.float32(s0))
        c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A   # interpolateCubic
        c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
        y = one - x
        c2 = ((A + np.float32(2)) * y - (A + np.float32(3))) * y * y + one
        c3 = one - c0 - c1 - c2
        taps[d] = np.rint(np.array([c0, c1, c2, c3], np.float32) * np.float32(2048)).astype(np.int64)     # saturate_cast<short>
        ofs[d] = s0
    return ofs, taps


def resize_cubic_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC) on 8-bit data, OpenCV's own kernel (no IPP): taps at
    offsets -1..2 with replicated borders, HResizeCubic in int32, VResizeCubic with FixedPtCast<int, uchar, 22>."""
    sh, sw, _ = img.shape
    xo, xa = _cubic_tab(dw, sw)
    yo, ya = _cubic_tab(dh, sh)
    S = img.astype(np.int64)
    xi = np.clip(xo[:, None] + np.arange(-1, 3)[None, :], 0, sw - 1)
    yi = np.clip(yo[:, None] + np.arange(-1, 3)[None, :], 0, sh - 1)
    H = (S[:, xi, :] * xa[None, :, :, None]).sum(2)                               # [sh, dw, 3]
    rows = H[yi, :, :].reshape(dh, 4, dw * 3)                                     # the four source rows of every output row
    V = (rows * ya[:, :, None]).sum(1)
    out = np.clip((V + (1 << 21)) >> 22, 0, 255).astype(np.uint8)                 # scalar tail: FixedPtCast<int, uchar, 22>
    # VResizeCubicVec_32s8u (128-bit universal intrinsics of the baseline build: eight elements per step, no fused
    # multiply-add): float32 evaluation S3*b3, then S2*b2 + t, S1*b1 + t, S0*b0 + t with b_k = beta_k / 2^22,
    # round half to even.  Differs from the integer rounding on ~1e-5 of the elements.
    nvec = (dw * 3 // 8) * 8
    if nvec:
        f = rows[:, :, :nvec].astype(np.float32)
        bf = (ya.astype(np.float32) * np.float32(1.0 / (2048 * 2048)))[:, :, None]
        t = f[:, 3] * bf[:, 3]
        for k in (2, 1, 0):
            t = f[:, k] * bf[:, k] + t
        out[:, :nvec] = np.clip(np.rint(t), 0, 255).astype(np.uint8)
    return out.reshape(dh, dw, 3)


def print_patch(patch_u8: np.ndarray) -> np.ndarray:
    """adv_patch.py:40-59: deterministic print adjust, float64, truncating cast."""
    p = patch_u8 - 127.0
    p /= 128.0
    p *= 0.5
    p *= 128.0
    p += 127.0
    return np.clip(p, 0.0, 255.0).astype(np.uint8)


def create(img_h: int, img_w: int, bbox, scale: float):
    """adv_patch.py:61-92 -> (ymin_patch, xmin_patch, patch_h, patch_w) ints."""
    ymin, xmin, ymax, xmax = bbox
    h, w = ymax - ymin, xmax - xmin
    patch_w = int(max(h, w) * scale)
    patch_h = patch_w
    ymin_patch = max(ymin + h / 2.0 - patch_h / 2.0, 0.0)
    xmin_patch = max(xmin + w / 2.0 - patch_w / 2.0, 0.0)
    if ymin_patch + patch_h > img_h:
        ymin_patch = img_h - patch_h
    if xmin_patch + patch_w > img_w:
        xmin_patch = img_w - patch_w
    return int(ymin_patch), int(xmin_patch), patch_h, patch_w


def rescaled_target_ysum(frame: np.ndarray, output_size):
    """Sum of the Y channel of `rescale(frame)` (adv_patch.py:94-112) -> (integer sum, pixel count)."""
    h, w, _ = frame.shape
    image_scale = min(output_size[1] / w, output_size[0] / h)
    sh, sw = int(h * image_scale), int(w * image_scale)
    canvas = np.full((output_size[0], output_size[1], 3), 127, np.uint8)
    canvas[:sh, :sw] = resize_linear_u8(frame, sw, sh)
    return int(rgb2yuv(canvas)[..., 0].astype(np.int64).sum()), output_size[0] * output_size[1]


def transformed_patch(frame: np.ndarray, patch_printed: np.ndarray, output_size, ph: int, pw: int, noise: np.ndarray):
    """adv_patch.py:162-177 with the np.random.uniform(-.01,.01) draw as the explicit float64 input `noise` [ph,pw,3]."""
    tsum, tn = rescaled_target_ysum(frame, output_size)
    src = rgb2yuv(patch_printed)
    source_mean = np.float64(src[..., 0].astype(np.int64).sum()) / np.float64(src[..., 0].size)
    target_mean = np.float64(tsum) / np.float64(tn)
    res = np.clip(src[..., 0] - source_mean + target_mean, 0.0, 255.0)
    src[..., 0] = res.astype(np.uint8)
    patch = yuv2rgb(src)
    h = patch.shape[0]
    if h > ph:
        patch = resize_area_u8(patch, pw, ph)
    elif h < ph:
        patch = resize_cubic_u8(patch, pw, ph)
    p = patch - 127.0
    p /= 128.0
    p = np.clip(p + noise, -1.0, 1.0)
    p *= 128.0
    p += 127.0
    return np.clip(p, 0.0, 255.0).astype(np.uint8)


def add_adv_to_img(frame: np.ndarray, bboxes, patch_printed: np.ndarray, output_size, scale: float, noises):
    """adv_patch.py:179-190: boxes are pasted in order; each brightness match sees the earlier pastes."""
    img = frame.copy()
    for bbox, noise in zip(bboxes, noises):
        y, x, ph, pw = create(img.shape[0], img.shape[1], bbox, scale)
        img[y:y + ph, x:x + pw] = transformed_patch(img, patch_printed, output_size, ph, pw, noise)
    return img
