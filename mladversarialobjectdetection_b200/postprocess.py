"""First-pass post-processing: person candidates -> NMS -> clipped ragged boxes
(reference: attacker.py:100-116, 143-170; tf2/postprocess.py:159-205 -> tf.raw_ops.NonMaxSuppressionV5).

This is the caller side of the hot path.  The candidate scores come from the score
kernel's workspace (score_max_fwd already applied arg-max-class == person and the valid-box filter); decode,
per-image (soft-)NMS in TF's lazy priority-queue order, clip_boxes and the CSR assembly all run in libeotpatch.so
(`person_nms`, csrc/nms.cu) -- the reference runs them as a host-synchronous map_fn over the images with TF's
CPU-only NMS kernel.  The only host read is row_splits (B+1 int32): the box count sizes the patcher's launch.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .ragged import RaggedBoxes


def nms_settings(nms_configs: dict):
    """tf2/postprocess.py:175-199 -> (soft_nms_sigma passed to TF, iou_threshold, score_threshold)."""
    method = nms_configs.get("method")
    if method == "hard" or not method:
        return 0.0, nms_configs.get("iou_thresh") or 0.5, nms_configs.get("score_thresh") or float("-inf")
    if method == "gaussian":
        sigma = nms_configs.get("sigma") or 0.5
        return sigma / 2, 1.0, nms_configs.get("score_thresh") or 0.001
    raise ValueError("Inference has invalid nms method {}".format(method))


def person_boxes_after_nms(config, score_ctx, box_outputs: Sequence[torch.Tensor], anchors: torch.Tensor,
                           image_hw, thresh: bool = True, max_candidates: int = 0, box_capacity=None):
    """attacker.py:104-116: candidates (person & valid, from the score kernel) [>= score_thresh] -> NMS ->
    clip -> ragged boxes / scores.

    box_capacity None: one small device->host read (row_splits) sizes the ragged result exactly and the per-image
    score lists are returned (metrics / ASR).  box_capacity = n: NO host read -- the boxes come back at that capacity
    (the kernels downstream read the count in use, row_splits[-1], on the device) and the scores stay on the device
    (second return value: the NmsResult)."""
    sigma, iou_thresh, nms_score_thresh = nms_settings(config.nms_configs)
    floor = float(config.nms_configs["score_thresh"]) if thresh else 0.0
    res = ops.person_nms(ops.score_candidate_view(score_ctx), box_outputs, anchors, image_hw,
                         max_output_size=int(config.nms_configs["max_output_size"]), iou_threshold=float(iou_thresh),
                         score_threshold=max(float(nms_score_thresh), 0.0), soft_nms_sigma=float(sigma),
                         score_floor=max(floor, 0.0), max_candidates=max_candidates)
    if box_capacity is not None:
        cap = min(int(box_capacity), int(res.ragged_boxes.shape[0]))
        return RaggedBoxes(res.ragged_boxes[:cap], res.row_splits), res
    splits = res.row_splits.cpu()                                             # the one host sync
    n = int(splits[-1])
    if n < 0:
        raise RuntimeError("person_nms: an image holds more candidates than max_candidates; raise it (0 = all anchors)")
    boxes = RaggedBoxes(res.ragged_boxes[:n], res.row_splits)
    s = splits.numpy()
    scores = res.ragged_scores[:n].cpu().numpy()
    return boxes, [scores[s[i]:s[i + 1]] for i in range(len(s) - 1)]


def calc_asr(scores: Sequence[np.ndarray], scores_pred: Sequence[np.ndarray], score_thresh: float = 0.5) -> float:
    """attack success rate at a score threshold (attacker.py:238-255): 1 - (#attacked-pass boxes with score >= t) /
    (#clean-pass boxes with score >= t), both after NMS; `scores` / `scores_pred` are the ragged per-image score lists
    that `person_boxes_after_nms` returns for the clean and the attacked pass.  The reference counts the 4 coordinates
    of every kept box (`tf.size(flat_values)`) and adds Keras' epsilon (1e-7) to the denominator; float32 throughout."""
    t = np.float32(score_thresh)
    n_first = sum(int((np.asarray(s, np.float32) >= t).sum()) for s in scores)
    n_pred = sum(int((np.asarray(s, np.float32) >= t).sum()) for s in scores_pred)
    return float(np.float32(1.0) - np.float32(4 * n_pred) / (np.float32(4 * n_first) + np.float32(1e-7)))


def asr_sweep(scores, scores_pred, bins: Sequence[float]) -> np.ndarray:
    """the threshold sweep of vis_images (attacker.py:275-277) over `PatchAttacker.bins`."""
    return np.asarray([calc_asr(scores, scores_pred, float(b)) for b in bins], np.float32)
