"""The oracle against the REFERENCE'S OWN `attacker.Patcher.call` and `BrightnessMatcher.call`, executed verbatim in
the build container on a NumPy stand-in for the TF ops they call (tests/golden/tf_numpy_shim.py; fixtures
tests/golden/patcher_ref.npz / brightness_ref.npz written by tests/golden/make_golden.py).

What this pins: the reference's Python -- expression order / association, casts, pad split, where / clip / scatter
sequence, the area filter (a dropped box sits in the middle of image 0's list), the loop over boxes and images, the
brightness matcher's formula order.  The leaf TF kernels underneath are the oracle's own restatements (unpinned)."""
import os

import numpy as np
import pytest

from oracle import patcher
from oracle.patcher import BOX_PARAMS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_patcher_fixture():
    g = np.load(os.path.join(GOLD, "patcher_ref.npz"))
    params = np.ascontiguousarray(g["params"]).view(BOX_PARAMS).reshape(-1)
    off = g["offsets"]
    boxes = [g["boxes"][off[b]:off[b + 1]] for b in range(len(off) - 1)]
    prm = [params[off[b]:off[b + 1]] for b in range(len(off) - 1)]
    return g, boxes, prm


def test_oracle_equals_reference_patcher_on_the_shim():
    g, boxes, prm = load_patcher_fixture()
    out, _, states = patcher.patcher_forward(g["patch"], g["images"], boxes, prm, g["print_wb"], float(g["scale"]))
    np.testing.assert_array_equal(out, g["out_ref"])
    assert len(boxes[0]) == 3 and len(states[0].boxes) == 2          # the area filter dropped the tiny box
    assert (out != g["images"]).any()


def test_oracle_equals_reference_brightness_matcher_on_the_shim():
    g = np.load(os.path.join(GOLD, "brightness_ref.npz"))
    np.testing.assert_array_equal(patcher.brightness_match(g["src"], g["tgt"]), g["out_ref"])


def load_masker_fixture(tag):
    g = np.load(os.path.join(GOLD, "masker_ref.npz"))
    params = np.ascontiguousarray(g[f"{tag}_params"]).view(BOX_PARAMS).reshape(-1)
    off = g[f"{tag}_offsets"]
    boxes = [g[f"{tag}_boxes"][off[b]:off[b + 1]] for b in range(len(off) - 1)]
    prm = [params[off[b]:off[b + 1]] for b in range(len(off) - 1)]
    return g, boxes, prm, params


def test_oracle_equals_reference_masker_on_the_shim():
    """attack_detection.Masker.call (attack_detection.py:321-498): evaluation branch (learned patch, tolerance 0) and
    training branch (shuffled / flipped crops of other batch images as per-image patches, scale U(.3,.5), noise +-0.1)."""
    for tag, tol in (("eval", 0.0), ("train", 0.5)):
        g, boxes, prm, _ = load_masker_fixture(tag)
        out, mask, _ = patcher.patcher_forward(g[f"{tag}_patch"], g[f"{tag}_images"], boxes, prm, g[f"{tag}_print_wb"],
                                               float(g[f"{tag}_scale"]), tolerance=tol, noise_amp=0.1, want_mask=True)
        np.testing.assert_array_equal(out, g[f"{tag}_out"])
        np.testing.assert_array_equal(mask, g[f"{tag}_mask"])
        assert mask.any()


def test_oracle_equals_reference_second_and_first_pass_on_the_shim():
    """PatchAttacker.second_pass / first_pass + the max-score line (attacker.py:69-170,190) with the real vendored
    pre_nms / nms / clip_boxes / Anchors underneath (fixtures objective_ref.npz)."""
    from oracle import nms as onms, objective
    from tests._util import objective_fixture_inputs
    g = np.load(os.path.join(GOLD, "objective_ref.npz"))
    B, H = int(g["B"]), int(g["H"])
    cls, box = objective_fixture_inputs(int(g["seed"]), B, H)
    anchors = objective.anchor_boxes(H)
    c_all, b_all = objective.merge_levels(cls, box, 90)
    post = objective.objective_forward(c_all, b_all, anchors, H, H, 0.4)
    np.testing.assert_array_equal(post["max_scores"], g["max_scores"])
    cand = np.where(post["cand"], post["score"], np.float32(-1.0)).astype(np.float32)
    cfg = dict(method="gaussian", sigma=None, iou_thresh=0.5, score_thresh=0.5, max_output_size=100)
    rows, row_scores = onms.person_boxes_after_nms(cand, b_all, anchors, (H, H), cfg, thresh=True)
    for b in range(B):
        np.testing.assert_array_equal(post["score"][b][post["cand"][b]], g[f"sp_scores{b}"])
        np.testing.assert_array_equal(post["boxes"][b][post["cand"][b]], g[f"sp_boxes{b}"])
        np.testing.assert_array_equal(rows[b], g[f"fp_boxes{b}"])
        np.testing.assert_array_equal(row_scores[b], g[f"fp_scores{b}"])
        assert len(g[f"fp_boxes{b}"]) > 5
    hard = dict(method="hard", sigma=None, iou_thresh=0.5, score_thresh=0.4, max_output_size=100)
    rows, row_scores = onms.person_boxes_after_nms(cand, b_all, anchors, (H, H), hard, thresh=True)
    for b in range(B):                                              # hard-NMS branch of postprocess.nms
        np.testing.assert_array_equal(rows[b], g[f"hard_boxes{b}"])
        np.testing.assert_array_equal(row_scores[b], g[f"hard_scores{b}"])


def test_oracle_equals_reference_patcher_on_the_shim_harder_case():
    """patcher_ref2: 3 images of 96x96, 32x32 patch, scale .6 -- patch sides from below to above the texture size (down-
    and up-sampling), mutually overlapping boxes, windows clamped at the image border, one box dropped by the area filter."""
    g = np.load(os.path.join(GOLD, "patcher_ref2.npz"))
    params = np.ascontiguousarray(g["params"]).view(BOX_PARAMS).reshape(-1)
    off = g["offsets"]
    boxes = [g["boxes"][off[b]:off[b + 1]] for b in range(len(off) - 1)]
    prm = [params[off[b]:off[b + 1]] for b in range(len(off) - 1)]
    out, _, states = patcher.patcher_forward(g["patch"], g["images"], boxes, prm, g["print_wb"], float(g["scale"]))
    np.testing.assert_array_equal(out, g["out_ref"])
    sizes = sorted(bs.plan.ps for st in states for bs in st.boxes)
    assert sizes[0] < 32 < sizes[-1] and sum(len(st.boxes) for st in states) == 8


def test_attack_success_rate_matches_reference_calc_asr():
    """postprocess.calc_asr / asr_sweep (host metric) vs the reference's own calc_asr (attacker.py:238-255) swept over
    PatchAttacker.bins, run on the NumPy TF shim (objective_ref.npz)."""
    from mladversarialobjectdetection_b200 import postprocess
    g = np.load(os.path.join(GOLD, "objective_ref.npz"))
    B = int(g["B"])
    first = [g[f"hard_scores{b}"] for b in range(B)]
    attacked = [g[f"atk_scores{b}"] for b in range(B)]
    np.testing.assert_array_equal(postprocess.asr_sweep(first, attacked, g["asr_bins"]), g["asr"])
    assert postprocess.calc_asr(first, first, 0.5) == pytest.approx(0.0, abs=1e-6)
    assert postprocess.calc_asr(first, [np.zeros(0, np.float32)] * B, 0.5) == 1.0
