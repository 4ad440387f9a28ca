"""Oracle vs committed fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np

from mladversarialobjectdetection_b200 import synth
from oracle import objective, patcher

F = np.float32
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_centre_clamp_against_reference_adv_patch_create():
    """adv_patch.AdversarialPatch._create (reference, adv_patch.py:61-92) == the oracle's centring /
    clamping with extent = patch size, no jitter (same formula as attacker.py:480-486)."""
    rows = np.load(os.path.join(GOLD, "adv_patch_create.npz"))["rows"]
    assert len(rows) == 123
    for H, W, scale, ymin, xmin, ymax, xmax, ry, rx, rph, rpw in rows:
        h, w = ymax - ymin, xmax - xmin
        ps = int(max(h, w) * scale)
        y0 = patcher.centre_clamp(F(ymin + h / 2.0), F(ps), F(H))
        x0 = patcher.centre_clamp(F(xmin + w / 2.0), F(ps), F(W))
        assert (int(y0), int(x0), ps, ps) == (int(ry), int(rx), int(rph), int(rpw))
    # the survey's probe: box (50,125,400,200) on 480x640 at scale .5 -> [137, 75, 175, 175]
    assert [int(v) for v in rows[0][7:]] == [137, 75, 175, 175]


def test_oracle_matches_committed_vectors():
    g = np.load(os.path.join(GOLD, "oracle_small.npz"))
    bt = synth.make_batch(2, 64, 64, max_boxes=3, min_boxes=2, seed=314)
    patch = synth.make_patch(16, seed=3)
    bx, pr = bt.ragged()
    out, _, states = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, 0.4)
    np.testing.assert_array_equal(out, g["out"])
    plans = np.array([[bs.plan.y0, bs.plan.x0, bs.plan.ps, bs.plan.d, bs.plan.pad_lo]
                      for st in states for bs in st.boxes], dtype=np.int32)
    np.testing.assert_array_equal(plans, g["plans"])
    G = np.random.default_rng(int(g["G_seed"])).normal(size=out.shape).astype(F)
    gp = patcher.patcher_backward(G, patch, bt.print_wb, states)
    np.testing.assert_allclose(gp, g["grad_patch"], rtol=1e-5, atol=1e-7)
    anchors = objective.anchor_boxes(64)
    np.testing.assert_array_equal(anchors[:18], g["anchors_head"])
    np.testing.assert_array_equal(anchors[-9:], g["anchors_tail"])


def test_anchor_tables_equal_reference_anchors_class():
    """tf2/anchors.py:Anchors (the reference's own NumPy arithmetic, fixtures anchors_ref.npz): whole-table SHA-256 and
    every 97th row, for D0 512, lite 640, D4 1024, lite0 320 (anchor_scale 3) and a non-square input."""
    import hashlib
    from mladversarialobjectdetection_b200.anchors import anchor_table
    g = np.load(os.path.join(GOLD, "anchors_ref.npz"))
    for k, (h, w, scale) in enumerate(g["cfgs"]):
        for tab in (anchor_table((int(h), int(w)), 3, 7, 3, (1.0, 2.0, 0.5), float(scale)),
                    objective.anchor_boxes((int(h), int(w)), anchor_scale=float(scale))):
            assert len(tab) == int(g[f"n{k}"])
            np.testing.assert_array_equal(tab[::97], g[f"rows{k}"])
            assert hashlib.sha256(np.ascontiguousarray(tab.astype(np.float32)).tobytes()).hexdigest() == str(g[f"sha{k}"])
