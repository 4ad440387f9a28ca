"""Synthetic COCO-shaped inputs and explicit transform seeds.

Every quantity the reference draws from TF's RNG inside `Patcher` (attacker.py:370-371,
426-427, 436, 473-474) is drawn here, per GLOBAL image index, so that a batch sharded
over G ranks sees exactly the transforms the single-GPU batch sees.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

F = np.float32

# C-ABI record `EotBoxParams` (include/eotpatch.h) -- 48 bytes.
BOX_PARAMS = np.dtype([("uy", "f4"), ("ux", "f4"), ("delta", "f4"), ("cos", "f4"), ("sin", "f4"),
                       ("pa", "f4"), ("pb", "f4"), ("scale", "f4"),
                       ("key0", "u4"), ("key1", "u4"), ("rsv0", "u4"), ("rsv1", "u4")])

MAX_ANGLE = 20.0 * np.pi / 180.0      # attacker.py:436
MAX_DELTA = 0.3                        # attacker.py:427


@dataclass
class Batch:
    images: np.ndarray          # [B,H,W,3] float32 in [-1,1]
    boxes: np.ndarray           # [N,4] float32 (ymin,xmin,ymax,xmax) pixels, ragged rows concatenated
    offsets: np.ndarray         # [B+1] int32 CSR row splits of `boxes`
    params: np.ndarray          # [N] BOX_PARAMS
    print_wb: np.ndarray        # [B,6] float32 (w0,w1,w2,b0,b1,b2)

    def boxes_of(self, b: int) -> np.ndarray:
        return self.boxes[self.offsets[b]:self.offsets[b + 1]]

    def params_of(self, b: int) -> np.ndarray:
        return self.params[self.offsets[b]:self.offsets[b + 1]]

    def ragged(self):
        B = len(self.offsets) - 1
        return [self.boxes_of(b) for b in range(B)], [self.params_of(b) for b in range(B)]


def make_patch(P: int, seed: int = 7) -> np.ndarray:
    """Random initial patch U(-1,1) (attacker.py:43)."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, size=(P, P, 3)).astype(F)


def draw_box_params(rng: np.random.Generator, n: int, *, perspective: float = 0.0,
                    scale_range=None, max_angle: float = MAX_ANGLE) -> np.ndarray:
    """Transform seeds for n boxes.  cos/sin are evaluated in float64 and rounded once."""
    p = np.zeros(n, dtype=BOX_PARAMS)
    p["uy"] = rng.random(n, dtype=F)
    p["ux"] = rng.random(n, dtype=F)
    p["delta"] = (rng.random(n, dtype=F) * F(2 * MAX_DELTA) + F(-MAX_DELTA)).astype(F)
    ang = (rng.random(n, dtype=F) * (F(max_angle) - F(-max_angle)) + F(-max_angle)).astype(F)
    p["cos"] = np.cos(ang.astype(np.float64)).astype(F)
    p["sin"] = np.sin(ang.astype(np.float64)).astype(F)
    if perspective > 0:
        p["pa"] = rng.uniform(-perspective, perspective, n).astype(F)
        p["pb"] = rng.uniform(-perspective, perspective, n).astype(F)
    if scale_range is None:
        p["scale"] = F(-1.0)
    else:   # Masker training: scale ~ U(.3,.5) per box (attack_detection.py:453)
        lo, hi = F(scale_range[0]), F(scale_range[1])
        p["scale"] = (rng.random(n, dtype=F) * (hi - lo) + lo).astype(F)
    keys = rng.integers(0, 2 ** 32, size=(n, 2), dtype=np.uint64)
    p["key0"] = keys[:, 0].astype(np.uint32)
    p["key1"] = keys[:, 1].astype(np.uint32)
    return p


def draw_print_wb(rng: np.random.Generator) -> np.ndarray:
    """w ~ N(.5,.1)^3, b ~ N(0,.01)^3 (attacker.py:370-371)."""
    w = rng.normal(0.5, 0.1, 3)
    b = rng.normal(0.0, 0.01, 3)
    return np.concatenate([w, b]).astype(F)


def make_batch(B: int, H: int, W: int, *, first_image: int = 0, seed: int = 1234, max_boxes: int = 8,
               min_boxes: int = 1, perspective: float = 0.0, scale_range=None,
               image_fill: str = "uniform") -> Batch:
    """Images [first_image, first_image+B) of the synthetic stream."""
    images = np.empty((B, H, W, 3), dtype=F)
    boxes, params, wbs, offs = [], [], [], [0]
    for i in range(B):
        g = first_image + i
        rng = np.random.default_rng([seed, g])
        if image_fill == "uniform":
            images[i] = rng.random((H, W, 3), dtype=F) * F(2.0) - F(1.0)
        else:   # smooth scene: cheaper to reason about in tests
            yy, xx = np.mgrid[0:H, 0:W].astype(F)
            base = np.stack([np.sin(xx / 37 + g), np.cos(yy / 23 - g), np.sin((xx + yy) / 51)], -1)
            images[i] = (0.8 * base).astype(F)
        n = int(rng.integers(min_boxes, max_boxes + 1)) if max_boxes > 0 else 0
        bh = rng.uniform(0.2 * H, 0.8 * H, n)
        bw = rng.uniform(0.1 * W, 0.4 * W, n)
        cy = rng.uniform(bh / 2, H - bh / 2)
        cx = rng.uniform(bw / 2, W - bw / 2)
        bx = np.stack([cy - bh / 2, cx - bw / 2, cy + bh / 2, cx + bw / 2], axis=1).astype(F)
        boxes.append(bx.reshape(-1, 4))
        params.append(draw_box_params(rng, n, perspective=perspective, scale_range=scale_range))
        wbs.append(draw_print_wb(rng))
        offs.append(offs[-1] + n)
    return Batch(images,
                 np.concatenate(boxes, axis=0).astype(F).reshape(-1, 4),
                 np.asarray(offs, dtype=np.int32),
                 np.concatenate(params) if params else np.zeros(0, dtype=BOX_PARAMS),
                 np.stack(wbs).astype(F))


def objective_inputs(seed: int = 8, B: int = 2, H: int = 64):
    """Seeded victim outputs for the objective / first-pass checks (the inputs of tests/golden/objective_ref.npz): class logits
    with ~5 % of the anchors pushed towards `person`, box regressions ~N(0, .3).  Returns (cls levels, box levels) as lists of
    float32 [B,h,w,810] / [B,h,w,36] arrays for an EfficientDet-D0-shaped head (levels 3..7, 9 anchors, 90 classes)."""
    from .anchors import feature_sizes
    rng = np.random.default_rng(seed)
    fs = feature_sizes((H, H), 7)[3:]
    cls = [rng.normal(-3, 2.5, (B, h, w, 810)).astype(np.float32) for h, w in fs]
    for c in cls:
        v = c.reshape(B, -1, 90)
        pick = rng.random(v.shape[:2]) < 0.05
        v[pick, 0] += 8.0
    box = [rng.normal(0, 0.3, (B, h, w, 36)).astype(np.float32) for h, w in fs]
    return cls, box
