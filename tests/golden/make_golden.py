"""Regenerates the committed fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

(a) `adv_patch_create.npz` -- outputs of the REFERENCE's own `adv_patch.AdversarialPatch._create`
    (/root/reference/adv_patch.py:61-92), the only importable piece of the patch-placement logic.
(b) `oracle_small.npz`     -- the oracle's forward/backward on a small seeded case, so that the
    oracle cannot drift silently and the GPU box (which has no /root/reference) can check both the
    oracle and the CUDA path against a committed vector.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def gen_adv_patch_create():
    sys.path.insert(0, "/root/reference")
    import adv_patch  # noqa: the reference module itself (NumPy + cv2 + PIL only)
    rng = np.random.default_rng(2022)
    rows = []
    for frame_hw, scale in [((480, 640), 0.5), ((640, 640), 0.4), ((1024, 1024), 0.25)]:
        ap = adv_patch.AdversarialPatch(scale=scale, h=32, w=32)
        img = np.zeros(frame_hw + (3,), np.uint8)
        boxes = [(50, 125, 400, 200)]
        for _ in range(40):
            h = int(rng.integers(8, frame_hw[0])); w = int(rng.integers(8, frame_hw[1]))
            y = int(rng.integers(0, frame_hw[0] - h + 1)); x = int(rng.integers(0, frame_hw[1] - w + 1))
            boxes.append((y, x, y + h, x + w))
        for bb in boxes:
            rows.append([frame_hw[0], frame_hw[1], scale, *bb, *ap._create(img, bb)])
    np.savez_compressed(os.path.join(HERE, "adv_patch_create.npz"), rows=np.asarray(rows, dtype=np.float64))


def gen_oracle_small():
    from mladversarialobjectdetection_b200 import synth
    from oracle import objective, patcher
    bt = synth.make_batch(2, 64, 64, max_boxes=3, min_boxes=2, seed=314)
    patch = synth.make_patch(16, seed=3)
    bx, pr = bt.ragged()
    out, _, states = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, 0.4)
    G = np.random.default_rng(15).normal(size=out.shape).astype(np.float32)
    gp = patcher.patcher_backward(G, patch, bt.print_wb, states)
    plans = np.array([[bs.plan.y0, bs.plan.x0, bs.plan.ps, bs.plan.d, bs.plan.pad_lo]
                      for st in states for bs in st.boxes], dtype=np.int32)
    rng = np.random.default_rng(16)
    fs = objective.feat_sizes(64, 7)[3:]
    cls = [rng.normal(-2, 2, (2, h, w, 810)).astype(np.float32) for h, w in fs]
    box = [rng.normal(0, 0.5, (2, h, w, 36)).astype(np.float32) for h, w in fs]
    anchors = objective.anchor_boxes(64)
    c, b = objective.merge_levels(cls, box, 90)
    post = objective.objective_forward(c, b, anchors, 64, 64, 0.4)
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), out=out, grad_patch=gp, plans=plans,
                        G_seed=15, max_scores=post["max_scores"], cand_count=post["cand"].sum(1),
                        anchors_head=anchors[:18], anchors_tail=anchors[-9:])


if __name__ == "__main__":
    gen_adv_patch_create()
    gen_oracle_small()
    print("fixtures written to", HERE)
