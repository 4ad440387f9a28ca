"""GPU twin of `adv_patch.AdversarialPatch` (csrc/adv_u8.cu) through the C ABI: bit-exact against the reference's own
outputs (tests/golden/adv_patch_u8.npz) and against the oracle on larger seeded frames."""
import os

import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200.adv_patch import AdversarialPatch
from oracle import adv_patch_u8 as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_matches_reference_fixtures_bit_for_bit():
    g = np.load(os.path.join(GOLD, "adv_patch_u8.npz"))
    P = int(g["P"])
    for k in range(int(g["n"])):
        boxes = [tuple(b) for b in g[f"boxes{k}"]]
        ap = AdversarialPatch(scale=float(g[f"scale{k}"]), h=P, w=P, patch=g[f"raw{k}"])
        np.testing.assert_array_equal(ap._patch_img.cpu().numpy(), g[f"printed{k}"])
        frame = g[f"frame{k}"]
        for bb in boxes:
            assert tuple(ap._create(frame, bb)) == o.create(frame.shape[0], frame.shape[1], bb, ap.scale)
        got = ap.add_adv_to_img(frame, boxes, noise=[g[f"noise{k}_{i}"] for i in range(len(boxes))])
        np.testing.assert_array_equal(got, g[f"result{k}"])


@pytest.mark.parametrize("frame_hw,P", [((480, 640), 640), ((720, 1280), 640), ((333, 517), 320), ((640, 640), 640)])
def test_video_sized_frames_against_oracle(frame_hw, P):
    """640-px patch on video frames: same-size, exact-2x and general letter-box; integer and general area ratios;
    overlapping boxes pasted in order."""
    rng = np.random.default_rng(P + frame_hw[0])
    raw = rng.integers(0, 256, size=(P, P, 3), dtype=np.uint8)
    frame = rng.integers(0, 256, size=frame_hw + (3,), dtype=np.uint8)
    H, W = frame_hw
    boxes = [(0.1 * H, 0.1 * W, 0.9 * H, 0.4 * W), (0.2 * H, 0.3 * W, 0.7 * H, 0.6 * W), (0.5 * H, 0.05 * W, 0.95 * H, 0.5 * W),
             (10, 20, 10 + 0.4 * P, 20 + 0.1 * P)]                      # last: long side * .5 == P/5 -> integer ratio 5
    ap = AdversarialPatch(scale=0.5, h=P, w=P, patch=raw)
    pl = ap.placements(H, W, boxes)
    noises = [rng.uniform(-0.01, 0.01, size=(int(p[2]), int(p[3]), 3)) for p in pl]
    got = ap.add_adv_to_img(frame, boxes, noise=noises)
    want = o.add_adv_to_img(frame, boxes, o.print_patch(raw), (P, P), 0.5, noises)
    np.testing.assert_array_equal(got, want)
    assert (got != frame).any(-1).sum() >= sum(int(p[2]) * int(p[3]) for p in pl) // 4


def test_device_frame_in_device_frame_out_and_no_boxes():
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, size=(64, 64, 3), dtype=np.uint8)
    ap = AdversarialPatch(scale=0.5, h=64, w=64, patch=raw, seed=3)
    frame = torch.from_numpy(rng.integers(0, 256, size=(100, 120, 3), dtype=np.uint8)).cuda()
    out = ap.add_adv_to_img(frame, [])
    assert out.is_cuda and torch.equal(out, frame) and out.data_ptr() != frame.data_ptr()
    out = ap.add_adv_to_img(frame, [(10, 10, 90, 60)])                 # device-drawn noise
    y, x, ph, pw = ap._create(frame, (10, 10, 90, 60))
    changed = (out != frame).any(-1)
    assert changed[y:y + ph, x:x + pw].float().mean() > 0.9 and not changed[:y].any() and not changed[:, :x].any()


def test_bicubic_upsampling_branch_against_oracle():
    """boxes that need the patch larger than its texture (adv_patch.py:158-160, INTER_CUBIC), mixed with down-sampled
    ones; also pinned by the last case of the reference fixture above."""
    rng = np.random.default_rng(33)
    for P, hw in [(32, (200, 240)), (100, (512, 512)), (48, (301, 277))]:
        raw = rng.integers(0, 256, size=(P, P, 3), dtype=np.uint8)
        frame = rng.integers(0, 256, size=hw + (3,), dtype=np.uint8)
        H, W = hw
        boxes = [(0, 0, 0.95 * H, 0.5 * W), (0.1 * H, 0.3 * W, 0.9 * H, 0.9 * W), (5, 5, 5 + P, 5 + 0.5 * P)]
        ap = AdversarialPatch(scale=0.5, h=P, w=P, patch=raw)
        pl = ap.placements(H, W, boxes)
        assert any(int(p[2]) > P for p in pl) and any(int(p[2]) < P for p in pl)
        noises = [rng.uniform(-0.01, 0.01, size=(int(p[2]), int(p[3]), 3)) for p in pl]
        got = ap.add_adv_to_img(frame, boxes, noise=noises)
        want = o.add_adv_to_img(frame, boxes, o.print_patch(raw), (P, P), 0.5, noises)
        np.testing.assert_array_equal(got, want)


def test_invalid_boxes_raise_before_any_launch():
    raw = np.zeros((32, 32, 3), np.uint8)
    ap = AdversarialPatch(scale=0.5, h=32, w=32, patch=raw)
    frame = np.zeros((200, 200, 3), np.uint8)
    with pytest.raises(RuntimeError, match="empty patch"):
        ap.add_adv_to_img(frame, [(5, 5, 6, 6)])


def test_non_square_patch_texture_and_non_square_letterbox():
    """h != w everywhere: 96x64 patch texture (different area ratios per axis), AdversarialPatch(h=80, w=120)."""
    rng = np.random.default_rng(21)
    raw = rng.integers(0, 256, size=(96, 64, 3), dtype=np.uint8)
    frame = rng.integers(0, 256, size=(150, 200, 3), dtype=np.uint8)
    boxes = [(10, 10, 130, 80), (40, 90, 120, 190)]
    ap = AdversarialPatch(scale=0.4, h=80, w=120, patch=raw)
    pl = ap.placements(150, 200, boxes)
    noises = [rng.uniform(-0.01, 0.01, size=(int(p[2]), int(p[3]), 3)) for p in pl]
    got = ap.add_adv_to_img(frame, boxes, noise=noises)
    want = o.add_adv_to_img(frame, boxes, o.print_patch(raw), (80, 120), 0.4, noises)
    np.testing.assert_array_equal(got, want)
