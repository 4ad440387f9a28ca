"""First-pass post-processing: person candidates -> NMS -> clipped ragged boxes
(reference: attacker.py:100-116, 143-170; tf2/postprocess.py:159-205 -> tf.raw_ops.NonMaxSuppressionV5).

This is the caller side of the hot path.  The candidate scores come from the score
kernel's workspace (score_max_fwd already applied arg-max-class == person and the valid-box filter); the
inherently sequential per-image soft-NMS over the (few) candidates above the score threshold runs on the host,
exactly as the reference runs it on TF's CPU-only NonMaxSuppressionV5 kernel.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .ragged import RaggedBoxes


def _iou(a: np.ndarray, b: np.ndarray) -> float:
    ymin_a, xmin_a, ymax_a, xmax_a = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
    ymin_b, xmin_b, ymax_b, xmax_b = min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])
    area_a = (ymax_a - ymin_a) * (xmax_a - xmin_a)
    area_b = (ymax_b - ymin_b) * (xmax_b - xmin_b)
    if area_a <= 0 or area_b <= 0:
        return 0.0
    ih = max(min(ymax_a, ymax_b) - max(ymin_a, ymin_b), 0.0)
    iw = max(min(xmax_a, xmax_b) - max(xmin_a, xmin_b), 0.0)
    inter = ih * iw
    return float(inter / (area_a + area_b - inter))


def non_max_suppression_v5(boxes: np.ndarray, scores: np.ndarray, max_output_size: int, iou_threshold: float,
                           score_threshold: float, soft_nms_sigma: float):
    """tf.raw_ops.NonMaxSuppressionV5 (tensorflow/core/kernels/image/non_max_suppression_op.cc,
    DoNonMaxSuppressionOp) -> (selected indices, selected scores)."""
    import heapq
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    heap = [(-float(s), i, 0) for i, s in enumerate(scores) if s > score_threshold]
    heapq.heapify(heap)
    soft = soft_nms_sigma > 0
    scale = np.float32(-0.5 / soft_nms_sigma) if soft else np.float32(0.0)
    sel: List[int] = []
    sel_scores: List[float] = []
    while len(sel) < max_output_size and heap:
        neg, i, begin = heapq.heappop(heap)
        score = np.float32(-neg)
        original = score
        hard = False
        for j in range(len(sel) - 1, begin - 1, -1):
            sim = np.float32(_iou(boxes[i], boxes[sel[j]]))
            w = np.float32(np.exp(scale * sim * sim)) if sim <= iou_threshold else np.float32(0.0)
            score = np.float32(score * w)
            if not soft and sim > iou_threshold:
                hard = True
                break
            if score <= score_threshold:
                break
        if hard:
            continue
        if score == original:
            sel.append(i)
            sel_scores.append(float(score))
        elif score > score_threshold:
            heapq.heappush(heap, (-float(score), i, len(sel)))
    return np.asarray(sel, np.int64), np.asarray(sel_scores, np.float32)


def nms_settings(nms_configs: dict):
    """tf2/postprocess.py:175-199."""
    method = nms_configs.get("method")
    if method == "hard" or not method:
        return 0.0, nms_configs.get("iou_thresh") or 0.5, nms_configs.get("score_thresh") or float("-inf")
    if method == "gaussian":
        sigma = nms_configs.get("sigma") or 0.5
        return sigma / 2, 1.0, nms_configs.get("score_thresh") or 0.001
    raise ValueError("Inference has invalid nms method {}".format(method))


def decode_boxes(tb: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """decode_box_outputs (tf2/anchors.py:44-58) for gathered rows."""
    yca = (anchors[:, 0] + anchors[:, 2]) / 2
    xca = (anchors[:, 1] + anchors[:, 3]) / 2
    ha = anchors[:, 2] - anchors[:, 0]
    wa = anchors[:, 3] - anchors[:, 1]
    w = torch.exp(tb[:, 3]) * wa
    h = torch.exp(tb[:, 2]) * ha
    yc = tb[:, 0] * ha + yca
    xc = tb[:, 1] * wa + xca
    return torch.stack([yc - h / 2, xc - w / 2, yc + h / 2, xc + w / 2], dim=1)


def person_boxes_after_nms(config, score_ctx, box_outputs: Sequence[torch.Tensor], anchors: torch.Tensor,
                           image_hw, thresh: bool = True) -> Tuple[RaggedBoxes, List[np.ndarray]]:
    """attacker.py:104-116: candidates (person & valid, from the score kernel) [>= score_thresh] -> NMS ->
    clip -> ragged boxes/scores.  Synchronises (the result is ragged), like the reference's map_fn + NMS."""
    B = score_ctx.shape.batch
    A = score_ctx.shape.total_anchors
    from .ops import score_candidate_view
    cand = score_candidate_view(score_ctx)                                    # [B,A] float32, -1 = not a candidate
    sigma, iou_thresh, nms_score_thresh = nms_settings(config.nms_configs)
    floor = float(config.nms_configs["score_thresh"]) if thresh else 0.0
    keep = cand >= max(floor, 0.0)
    idx = keep.nonzero()                                                      # host sync (dynamic shape)
    rows, row_scores = [], []
    device = cand.device
    if idx.numel():
        tb_all = torch.cat([b.reshape(B, -1, 4) for b in box_outputs], dim=1)
        tb = tb_all[idx[:, 0], idx[:, 1]]
        boxes = decode_boxes(tb, anchors[idx[:, 1]]).cpu().numpy()
        scores = cand[idx[:, 0], idx[:, 1]].cpu().numpy()
        img = idx[:, 0].cpu().numpy()
    H, W = int(image_hw[0]), int(image_hw[1])
    for b in range(B):
        if idx.numel():
            m = img == b
            bb, ss = boxes[m], scores[m]
        else:
            bb, ss = np.zeros((0, 4), np.float32), np.zeros((0,), np.float32)
        sel, sel_scores = non_max_suppression_v5(bb, ss, config.nms_configs["max_output_size"], iou_thresh,
                                                 nms_score_thresh, sigma)
        out = bb[sel] if len(sel) else np.zeros((0, 4), np.float32)
        out = np.clip(out, 0, np.array([H, W, H, W], np.float32))             # postprocess.clip_boxes
        rows.append(out.astype(np.float32))
        row_scores.append(sel_scores)
    return RaggedBoxes.from_rows(rows, device), row_scores
