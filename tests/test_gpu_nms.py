"""person_nms (csrc/nms.cu) through the C ABI against the oracle's NonMaxSuppressionV5 restatement
(attacker.py:100-116,143-170; tf2/postprocess.py:159-205).  Bar: selected boxes and soft-NMS scores bit-exact."""
import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import anchors as anchors_mod, ops
from mladversarialobjectdetection_b200.anchors import feature_sizes
from oracle import nms as onms

pytestmark = pytest.mark.gpu
F = np.float32


def _case(B, H, seed, n_clusters=6, per_cluster=40, score_lo=0.3, empty=(), quantise=False):
    rng = np.random.default_rng(seed)
    fs = feature_sizes((H, H), 7)[3:]
    anchors = anchors_mod.anchor_table((H, H)).astype(F)
    A = anchors.shape[0]
    box_levels = [rng.normal(0, 0.25, (B, h, w, 36)).astype(F) for h, w in fs]
    cand = np.full((B, A), -1.0, F)
    for b in range(B):
        if b in empty:
            continue
        for _ in range(n_clusters):
            a0 = int(rng.integers(0, A - per_cluster))          # neighbouring anchors overlap heavily
            idx = a0 + rng.choice(per_cluster * 2, size=per_cluster, replace=False)
            idx = idx[idx < A]
            s = rng.uniform(score_lo, 0.99, len(idx)).astype(F)
            if quantise:
                s = (np.round(s * 20) / 20).astype(F)           # many exact score ties
            cand[b, idx] = s
    return cand, box_levels, anchors


def _run_and_compare(cand, box_levels, anchors, H, cfg, thresh=True, max_candidates=0):
    B = cand.shape[0]
    sigma, iou_thr, nms_thr = onms.nms_settings(cfg)
    floor = float(cfg["score_thresh"]) if thresh else 0.0
    dev = "cuda"
    res = ops.person_nms(torch.from_numpy(cand).to(dev), [torch.from_numpy(b).to(dev) for b in box_levels],
                         torch.from_numpy(anchors).to(dev), (H, H), max_output_size=cfg["max_output_size"],
                         iou_threshold=float(iou_thr), score_threshold=max(float(nms_thr), 0.0),
                         soft_nms_sigma=float(sigma), score_floor=max(floor, 0.0), max_candidates=max_candidates)
    reg = np.concatenate([b.reshape(B, -1, 4) for b in box_levels], 1)
    want_rows, want_scores = onms.person_boxes_after_nms(cand, reg, anchors, (H, H), cfg, thresh=thresh)
    splits = res.row_splits.cpu().numpy()
    vlen = res.valid_len.cpu().numpy()
    got_boxes = res.ragged_boxes.cpu().numpy()
    got_scores = res.ragged_scores.cpu().numpy()
    padded = res.nms_boxes.cpu().numpy()
    assert splits[0] == 0
    for b in range(B):
        n = len(want_rows[b])
        assert vlen[b] == n and splits[b + 1] - splits[b] == n
        np.testing.assert_array_equal(got_boxes[splits[b]:splits[b + 1]], want_rows[b])
        np.testing.assert_array_equal(got_scores[splits[b]:splits[b + 1]], want_scores[b])
        np.testing.assert_array_equal(padded[b, :n], want_rows[b])
        assert not padded[b, n:].any()
    return want_rows


GAUSS = dict(method="gaussian", sigma=None, iou_thresh=0.5, score_thresh=0.5, max_output_size=100)   # attacker_train.py:31


def test_gaussian_soft_nms_matches_oracle_incl_empty_images():
    cand, box_levels, anchors = _case(5, 256, seed=1, empty=(2,))
    rows = _run_and_compare(cand, box_levels, anchors, 256, GAUSS)
    assert len(rows[2]) == 0 and sum(len(r) for r in rows) > 50


def test_hard_nms_matches_oracle():
    cand, box_levels, anchors = _case(4, 256, seed=2)
    cfg = dict(method="hard", sigma=None, iou_thresh=0.5, score_thresh=0.4, max_output_size=100)
    _run_and_compare(cand, box_levels, anchors, 256, cfg)


def test_exact_score_ties_lower_index_first():
    cand, box_levels, anchors = _case(3, 256, seed=3, quantise=True)
    _run_and_compare(cand, box_levels, anchors, 256, GAUSS)


def test_more_than_1024_candidates_and_output_cap():
    """global-memory candidate path (> 1024 per image) and max_output_size reached"""
    cand, box_levels, anchors = _case(2, 512, seed=4, n_clusters=40, per_cluster=60, score_lo=0.55)
    assert (cand[0] >= 0.5).sum() > 1024
    cfg = dict(GAUSS, max_output_size=30)
    rows = _run_and_compare(cand, box_levels, anchors, 512, cfg)
    assert len(rows[0]) == 30


def test_no_threshold_second_pass_style_and_low_nms_threshold():
    cand, box_levels, anchors = _case(2, 128, seed=5, n_clusters=3, per_cluster=20, score_lo=0.01)
    cfg = dict(method="gaussian", sigma=0.3, iou_thresh=None, score_thresh=0.0, max_output_size=100)   # -> 0.001
    _run_and_compare(cand, box_levels, anchors, 128, cfg, thresh=False)


def test_candidate_overflow_is_reported_not_truncated():
    cand, box_levels, anchors = _case(2, 128, seed=6, n_clusters=4, per_cluster=30)
    dev = "cuda"
    res = ops.person_nms(torch.from_numpy(cand).to(dev), [torch.from_numpy(b).to(dev) for b in box_levels],
                         torch.from_numpy(anchors).to(dev), (128, 128), max_candidates=16)
    assert int(res.row_splits[-1]) == -1 and (res.valid_len.cpu().numpy() == -1).any()


def test_first_pass_emits_ragged_boxes_for_the_patcher():
    """PatchAttacker.first_pass (attacker.py:91-116) end to end on the device: score kernel -> person_nms -> RaggedBoxes."""
    from mladversarialobjectdetection_b200 import postprocess
    from oracle import objective
    rng = np.random.default_rng(8)
    B, H = 3, 128
    fs = feature_sizes((H, H), 7)[3:]
    cls = [rng.normal(-3, 2.5, (B, h, w, 810)).astype(F) for h, w in fs]
    for c in cls:                                          # make persons likely: boost class 0 of some anchors
        v = c.reshape(B, -1, 90)
        pick = rng.random(v.shape[:2]) < 0.05
        v[pick, 0] += 8.0
    box = [rng.normal(0, 0.3, (B, h, w, 36)).astype(F) for h, w in fs]
    anchors = anchors_mod.anchor_table((H, H)).astype(F)
    dev = "cuda"
    tc, tb = [torch.from_numpy(c).to(dev) for c in cls], [torch.from_numpy(b).to(dev) for b in box]
    ta = torch.from_numpy(anchors).to(dev)
    _, _, _, ctx = ops.score_max_forward(tc, tb, ta, (H, H))

    class Cfg:
        nms_configs = GAUSS
    got, got_scores = postprocess.person_boxes_after_nms(Cfg, ctx, tb, ta, (H, H), thresh=True)
    cand = ops.score_candidate_view(ctx).cpu().numpy()
    reg = np.concatenate([b.reshape(B, -1, 4) for b in box], 1)
    want_rows, want_scores = onms.person_boxes_after_nms(cand, reg, anchors, (H, H), GAUSS, thresh=True)
    rows = got.to_rows()
    assert sum(len(r) for r in want_rows) > 10
    for b in range(B):
        np.testing.assert_array_equal(rows[b], want_rows[b])
        np.testing.assert_array_equal(got_scores[b], want_scores[b])


def test_score_kernel_and_nms_equal_reference_passes_fixture():
    """score_max_fwd + person_nms vs the reference's own second_pass / first_pass run on the NumPy TF shim
    (tests/golden/objective_ref.npz): per-image max score, candidate scores, first-pass boxes and soft-NMS scores."""
    import os
    from mladversarialobjectdetection_b200 import postprocess
    from tests._util import objective_fixture_inputs
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "objective_ref.npz"))
    B, H = int(g["B"]), int(g["H"])
    cls, box = objective_fixture_inputs(int(g["seed"]), B, H)
    dev = "cuda"
    tc, tb = [torch.from_numpy(c).to(dev) for c in cls], [torch.from_numpy(b).to(dev) for b in box]
    ta = torch.from_numpy(anchors_mod.anchor_table((H, H)).astype(F)).to(dev)
    M, _, ncand, ctx = ops.score_max_forward(tc, tb, ta, (H, H))
    np.testing.assert_array_equal(M.cpu().numpy(), g["max_scores"])
    cand = ops.score_candidate_view(ctx).cpu().numpy()
    for b in range(B):
        np.testing.assert_array_equal(cand[b][cand[b] >= 0], g[f"sp_scores{b}"])
        assert int(ncand[b]) == len(g[f"sp_scores{b}"])

    class Cfg:
        nms_configs = GAUSS
    boxes, scores = postprocess.person_boxes_after_nms(Cfg, ctx, tb, ta, (H, H), thresh=True)
    rows = boxes.to_rows()
    for b in range(B):
        np.testing.assert_array_equal(rows[b], g[f"fp_boxes{b}"])
        np.testing.assert_array_equal(scores[b], g[f"fp_scores{b}"])
