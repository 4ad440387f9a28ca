"""Oracle NMS (oracle/nms.py: NonMaxSuppressionV5 restated) against the REFERENCE's own NumPy NMS
(automl/efficientdet/nms_np.py, fixtures tests/golden/nms_np.npz) and against hand-checkable cases."""
import os

import numpy as np

from oracle import nms

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_hard_nms_selects_what_reference_nms_np_selects():
    g = np.load(os.path.join(GOLD, "nms_np.npz"))
    for k in range(int(g["n_cases"])):
        dets, want = g[f"dets{k}"], g[f"hard{k}"]
        sel, _ = nms.non_max_suppression_v5(dets[:, :4], dets[:, 4], 10000, 0.5, 0.0, 0.0)
        np.testing.assert_array_equal(dets[sel, :4].astype(np.float32), want[:, :4].astype(np.float32))


def test_gaussian_soft_nms_matches_reference_nms_np():
    """nms_np.soft_nms is the eager form (every remaining score decays at once); TF's kernel is lazy.  Same
    selection order, scores equal up to the +1 pixel-area convention of nms_np (boxes are thousands of units wide)."""
    g = np.load(os.path.join(GOLD, "nms_np.npz"))
    for k in range(int(g["n_cases"])):
        dets, want = g[f"dets{k}"], g[f"soft{k}"]
        # efficientdet passes sigma/2 to TF (tf2/postprocess.py:196-199): exp(-0.5/(sigma/2) iou^2) == exp(-iou^2/sigma)
        sel, scores = nms.non_max_suppression_v5(dets[:, :4], dets[:, 4], 10000, 1.0, 0.2, 0.25)
        np.testing.assert_array_equal(dets[sel, :4].astype(np.float32), want[:, :4].astype(np.float32))
        np.testing.assert_allclose(scores, want[:, 4], rtol=5e-3)


def test_score_ties_go_to_the_lower_index_and_threshold_is_strict():
    boxes = np.array([[0, 0, 10, 10], [100, 100, 110, 110], [200, 200, 210, 210]], np.float32)
    sel, _ = nms.non_max_suppression_v5(boxes, np.array([0.7, 0.9, 0.9], np.float32), 10, 0.5, 0.7, 0.0)
    assert sel.tolist() == [1, 2]                      # 0.7 > 0.7 is false: not admitted
    sel, sc = nms.non_max_suppression_v5(boxes[[0, 0, 1]], np.array([0.9, 0.8, 0.6], np.float32), 10, 0.5, 0.0, 0.0)
    assert sel.tolist() == [0, 2]                      # identical box suppressed (hard)
    sel, sc = nms.non_max_suppression_v5(boxes[[0, 0, 1]], np.array([0.9, 0.8, 0.6], np.float32), 10, 1.0, 0.0, 0.25)
    # Gaussian: IoU 1 -> weight exp(-2): 0.8 * exp(-2) = 0.108 -> selected last with the decayed score
    assert sel.tolist() == [0, 2, 1]
    np.testing.assert_allclose(sc, [0.9, 0.6, 0.8 * np.exp(-2.0)], rtol=1e-6)


def test_max_output_size_and_empty_input():
    boxes = np.stack([np.array([i * 20, 0, i * 20 + 10, 10], np.float32) for i in range(8)])
    sel, _ = nms.non_max_suppression_v5(boxes, np.linspace(0.9, 0.2, 8).astype(np.float32), 3, 0.5, 0.0, 0.0)
    assert sel.tolist() == [0, 1, 2]
    sel, sc = nms.non_max_suppression_v5(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 3, 0.5, 0.0, 0.0)
    assert len(sel) == 0 and len(sc) == 0


def test_nms_settings_follow_postprocess_nms():
    assert nms.nms_settings(dict(method="gaussian", sigma=None, iou_thresh=0.5, score_thresh=0.5)) == (0.25, 1.0, 0.5)
    assert nms.nms_settings(dict(method="hard", sigma=None, iou_thresh=None, score_thresh=0.0)) == (0.0, 0.5, float("-inf"))
