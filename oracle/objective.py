"""Oracle restatement of the person-score objective and its gradient.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * `PatchAttacker.second_pass` / `filter_valid_boxes`   /root/reference/attacker.py:69-89,118-141
  * `postprocess.pre_nms` (max-reduce branch)            automl/efficientdet/tf2/postprocess.py:67-79,104-116,136-156
  * `anchors.Anchors` / `decode_box_outputs`             automl/efficientdet/tf2/anchors.py:30-58,117-165
  * `utils.get_feat_sizes`                               automl/efficientdet/utils.py:509-526
  * the loss                                             attacker.py:190-193
Pinned: first normalised anchor [0.125,0.125,0.25,0.25] of tf2/postprocess_test.py:27-35,229.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import tfops
from .tfops import F


def feat_sizes(image_size, max_level: int):
    """Feature-map (h, w) for level 0..max_level: repeated (f-1)//2+1 (utils.py:520-526)."""
    if isinstance(image_size, int):
        image_size = (image_size, image_size)
    sizes = [tuple(image_size)]
    for _ in range(max_level):
        h, w = sizes[-1]
        sizes.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
    return sizes


def anchor_boxes(image_size, min_level=3, max_level=7, num_scales=3,
                 aspect_ratios=(1.0, 2.0, 0.5), anchor_scale=4.0) -> np.ndarray:
    """All anchors [A,4] (ymin,xmin,ymax,xmax) float32; order level -> y -> x -> (octave, aspect)."""
    if isinstance(image_size, int):
        image_size = (image_size, image_size)
    fs = feat_sizes(image_size, max_level)
    per_level = []
    for level in range(min_level, max_level + 1):
        sy = fs[0][0] / float(fs[level][0])
        sx = fs[0][1] / float(fs[level][1])
        ys = np.arange(sy / 2, image_size[0], sy)
        xs = np.arange(sx / 2, image_size[1], sx)
        cx, cy = np.meshgrid(xs, ys)
        cx = cx.reshape(-1)
        cy = cy.reshape(-1)
        shapes = []
        for octave in range(num_scales):
            for aspect in aspect_ratios:
                ax = np.sqrt(aspect)
                ay = 1.0 / ax
                hx = anchor_scale * sx * 2 ** (octave / float(num_scales)) * ax / 2.0
                hy = anchor_scale * sy * 2 ** (octave / float(num_scales)) * ay / 2.0
                shapes.append(np.stack([cy - hy, cx - hx, cy + hy, cx + hx], axis=1))
        per_level.append(np.stack(shapes, axis=1).reshape(-1, 4))      # [loc, 9, 4] -> [loc*9, 4]
    return np.concatenate(per_level, axis=0).astype(F)


def merge_levels(cls_levels: Sequence[np.ndarray], box_levels: Sequence[np.ndarray], num_classes: int):
    """merge_class_box_level_outputs (postprocess.py:67-79), channels_last."""
    B = cls_levels[0].shape[0]
    c = np.concatenate([x.reshape(B, -1, num_classes) for x in cls_levels], axis=1)
    b = np.concatenate([x.reshape(B, -1, 4) for x in box_levels], axis=1)
    return c.astype(F, copy=False), b.astype(F, copy=False)


def decode(tb: np.ndarray, anchors: np.ndarray) -> np.ndarray:
    """decode_box_outputs (anchors.py:44-58), float32."""
    a = anchors.astype(F)
    yca = (a[..., 0] + a[..., 2]) / F(2)
    xca = (a[..., 1] + a[..., 3]) / F(2)
    ha = a[..., 2] - a[..., 0]
    wa = a[..., 3] - a[..., 1]
    ty, tx, th, tw = tb[..., 0], tb[..., 1], tb[..., 2], tb[..., 3]
    with np.errstate(over="ignore"):
        w = np.exp(tw.astype(np.float64)).astype(F) * wa         # correctly rounded float32 exp (as oracle/nms.py)
        h = np.exp(th.astype(np.float64)).astype(F) * ha
    yc = ty * ha + yca
    xc = tx * wa + xca
    return np.stack([yc - h / F(2), xc - w / F(2), yc + h / F(2), xc + w / F(2)], axis=-1).astype(F)


def sigmoid(x: np.ndarray) -> np.ndarray:
    x64 = x.astype(np.float64)
    return (1.0 / (1.0 + np.exp(-x64))).astype(F)


def second_pass_post(cls_all: np.ndarray, box_all: np.ndarray, anchors: np.ndarray, H: int, W: int):
    """pre_nms + person filter + valid-box filter (attacker.py:130-140).

    Returns dict with per-anchor `cls`, `logit`, `score`, `boxes`, `cand` (bool [B,A])."""
    cls = np.argmax(cls_all, axis=-1).astype(np.int32)          # first max index wins ties
    logit = np.max(cls_all, axis=-1)
    with np.errstate(over="ignore", invalid="ignore"):           # exp() of a wild regression -> inf, as in TF
        boxes = decode(box_all, anchors[None])
        score = sigmoid(logit)
        bh = boxes[..., 2] - boxes[..., 0]
        bw = boxes[..., 3] - boxes[..., 1]
        area = bh * bw
        cond1 = (bw / F(W) <= F(1.0)) & (bh / F(H) <= F(1.0))
        cond2 = area > F(100.0)
    cand = (cls == 0) & cond1 & cond2
    return dict(cls=cls, logit=logit, score=score, boxes=boxes, cand=cand, area=area, bh=bh, bw=bw)


def objective_forward(cls_all, box_all, anchors, H, W, scale: float):
    """max_scores and the data part of the loss (attacker.py:190-191,193 without TV)."""
    post = second_pass_post(cls_all, box_all, anchors, H, W)
    sc = np.where(post["cand"], post["score"], F(-1.0))
    has = post["cand"].any(axis=1)
    max_scores = np.where(has, sc.max(axis=1), F(0.0)).astype(F)   # maximum(reduce_max, 0)
    scale = F(scale)
    scale_losses = (max_scores - scale) ** 2
    loss = F(np.sum((max_scores ** 2 + scale_losses).astype(np.float64)))
    post.update(max_scores=max_scores, has=has, loss=loss, scale_losses=scale_losses)
    return post


def objective_backward(cls_all, post, scale: float, dtype=F):
    """dL/dcls_all (dense, [B,A,C]) and dL/dscale for loss = sum(M^2 + (M-scale)^2).

    Ties split equally (UnsortedSegmentMax grad over anchors, Max grad over classes)."""
    B, A, C = cls_all.shape
    scale = F(scale)
    M = post["max_scores"]
    dM = (F(2) * M + F(2) * (M - scale)).astype(dtype)
    dscale = dtype(np.sum((-(F(2) * (M - scale))).astype(np.float64)))
    dcls = np.zeros((B, A, C), dtype=dtype)
    for b in range(B):
        if not post["has"][b]:
            continue
        sel = post["cand"][b] & (post["score"][b] == M[b])
        idx = np.nonzero(sel)[0]
        for a in idx:
            s = dtype(post["score"][b, a])
            dz = dM[b] / dtype(len(idx)) * s * (dtype(1) - s)
            ties = cls_all[b, a] == post["logit"][b, a]
            dcls[b, a, ties] = dz / dtype(ties.sum())
    return dcls, dscale


def split_levels(dcls_all: np.ndarray, level_shapes: Sequence[tuple]):
    """Inverse of merge_levels for the gradient: [B,A,C] -> list of [B,h,w,9*C]."""
    out, off = [], 0
    B = dcls_all.shape[0]
    for shp in level_shapes:
        n = int(np.prod(shp[1:])) // dcls_all.shape[2]
        out.append(dcls_all[:, off:off + n].reshape(shp))
        off += n
    return out


def attack_loss(max_scores: np.ndarray, scale: float, patch: np.ndarray):
    """loss, tv_loss, scale_loss (attacker.py:190-193)."""
    tv, tv_g = tfops.total_variation(patch)
    scale = F(scale)
    scale_losses = (max_scores - scale) ** 2
    loss = F(np.sum((max_scores ** 2 + scale_losses).astype(np.float64))) + F(1e-5) * tv
    return F(loss), tv, F(np.sum(scale_losses.astype(np.float64))), tv_g
