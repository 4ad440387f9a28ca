"""GPU parity: CUDA backward (dL/dpatch) vs the oracle's restated tape.gradient chain.
Bar (BASELINE.json north_star): patch gradients within 1e-4 relative L2."""
import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import ops, synth
from oracle import patcher
from tests._util import run_forward

pytestmark = pytest.mark.gpu
F = np.float32
REL_L2 = 1e-4


def _check_backward(bt, patch, scale, seed, smooth_grad=False, geom=None):
    out, _, ctx, d = run_forward(patch, scale, bt, geom)
    rng = np.random.default_rng(seed)
    G = rng.normal(size=bt.images.shape).astype(F)
    if smooth_grad:
        G = (G + np.roll(G, 1, 1) + np.roll(G, 1, 2) + np.roll(G, -1, 1) + np.roll(G, -1, 2)).astype(F)
    gp = ops.apply_backward(ctx, torch.from_numpy(G).cuda())
    torch.cuda.synchronize()
    bx, pr = bt.ragged()
    ref_out, _, states = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, scale)
    assert np.array_equal(out.cpu().numpy(), ref_out)
    ref = patcher.patcher_backward(G, patch, bt.print_wb, states, dtype=np.float64)
    got = gp.cpu().numpy().astype(np.float64)
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel <= REL_L2, f"rel-L2 {rel}"
    return gp, ctx, G, ref


@pytest.mark.parametrize("B,H,P,scale,seed", [
    (2, 64, 16, 0.4, 314),
    (4, 256, 100, 0.4, 21),
    (3, 320, 300, 0.4, 22),
    (2, 200, 640, 0.3, 23),
    (5, 128, 24, 0.9, 24),
    (20, 160, 40, 0.4, 26),      # more images than image groups of the texel pass
])
def test_backward_rel_l2(B, H, P, scale, seed):
    bt = synth.make_batch(B, H, H, seed=seed, max_boxes=6)
    _check_backward(bt, synth.make_patch(P, seed=seed), scale, seed)


def test_backward_serial_adjoint_path():
    # the memory-lean path (no per-box partial buffer) must give the same gradient
    bt = synth.make_batch(6, 192, 192, seed=27, max_boxes=5)
    patch = synth.make_patch(60, seed=27)
    g1, _, G, _ = _check_backward(bt, patch, 0.4, 27)
    g2, _, _, _ = _check_backward(bt, patch, 0.4, 27, geom=ops.PatchGeometry(serial_adjoint=True))
    assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-7 * float(g1.abs().max()))


def test_backward_dark_patch_clips_active():
    # a patch pushed against the clip limits exercises every mask of the chain
    bt = synth.make_batch(3, 192, 192, seed=61, max_boxes=4, min_boxes=2)
    patch = np.sign(synth.make_patch(48, seed=61)) * F(0.999)
    bt.print_wb[:, :3] = [1.4, 0.6, 2.2]
    _check_backward(bt, patch.astype(F), 0.4, 61)


def test_backward_overlapping_windows_route_to_last_paste():
    bt = synth.make_batch(2, 160, 160, seed=62, max_boxes=2, min_boxes=2)
    for b in range(2):      # force heavy overlap: same box twice
        bt.boxes[bt.offsets[b] + 1] = bt.boxes[bt.offsets[b]]
    _check_backward(bt, synth.make_patch(32, seed=62), 0.4, 62)


def test_backward_accumulate_and_zero_boxes():
    bt = synth.make_batch(2, 128, 128, seed=63, max_boxes=3)
    patch = synth.make_patch(32, seed=63)
    gp, ctx, G, _ = _check_backward(bt, patch, 0.4, 63)
    acc = torch.ones_like(gp)
    ops.apply_backward(ctx, torch.from_numpy(G).cuda(), grad_patch=acc, accumulate=True)
    torch.cuda.synchronize()
    assert torch.allclose(acc, gp + 1.0, rtol=0, atol=1e-6 * float(gp.abs().max()) + 1e-7)
    bt0 = synth.make_batch(2, 128, 128, seed=64, max_boxes=0)
    out, _, ctx0, _ = run_forward(patch, 0.4, bt0)
    g0 = ops.apply_backward(ctx0, torch.randn_like(out))
    torch.cuda.synchronize()
    assert float(g0.abs().max()) == 0.0


def test_backward_perspective_row():
    bt = synth.make_batch(3, 256, 256, seed=31, max_boxes=4, perspective=2e-4)
    _check_backward(bt, synth.make_patch(100, seed=31), 0.4, 31, smooth_grad=True)


def test_backward_config1_shape():
    bt = synth.make_batch(8, 512, 512, seed=1234, max_boxes=8)
    _check_backward(bt, synth.make_patch(100), 0.4, 7)


def test_tv_grad_and_adam_clip():
    from oracle import tfops
    patch = synth.make_patch(40, seed=70)
    p = torch.from_numpy(patch).cuda()
    g = torch.zeros_like(p)
    tv = ops.tv_grad_(p, g, 1e-5)
    torch.cuda.synchronize()
    tv_ref, g_ref = tfops.total_variation(patch)
    assert abs(float(tv) - float(tv_ref)) / float(tv_ref) < 1e-6
    np.testing.assert_allclose(g.cpu().numpy(), F(1e-5) * g_ref, rtol=0, atol=1e-12)
    # Adam (Keras ResourceApplyAdam) + clip, 3 steps, against a float64 restatement
    rng = np.random.default_rng(71)
    var = p.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    v64 = patch.astype(np.float64); m64 = np.zeros_like(v64); s64 = np.zeros_like(v64)
    for step in range(1, 4):
        gr = rng.normal(size=patch.shape).astype(F)
        ops.adam_clip_(var, m, v, torch.from_numpy(gr).cuda(), step, lr=1e-2, lo=-1.0, hi=1.0)
        alpha = 1e-2 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        m64 += (gr - m64) * (1 - 0.9); s64 += (gr.astype(np.float64) ** 2 - s64) * (1 - 0.999)
        v64 = np.clip(v64 - m64 * alpha / (np.sqrt(s64) + 1e-7), -1, 1)
    torch.cuda.synchronize()
    np.testing.assert_allclose(var.cpu().numpy(), v64, atol=2e-6)


def test_backward_many_mutually_overlapping_boxes():
    bt = synth.make_batch(2, 192, 192, seed=72, max_boxes=14, min_boxes=14)
    rng = np.random.default_rng(72)
    for j in range(bt.offsets[0], bt.offsets[1]):
        cy, cx = rng.uniform(70, 120, 2)
        h, w = rng.uniform(60, 130), rng.uniform(30, 70)
        bt.boxes[j] = [max(cy - h / 2, 0), max(cx - w / 2, 0), min(cy + h / 2, 192), min(cx + w / 2, 192)]
    _check_backward(bt, synth.make_patch(36, seed=72), 0.4, 72)


def test_backward_out_of_range_image():
    bt = synth.make_batch(3, 160, 160, seed=71, max_boxes=3, min_boxes=1)
    bt.images[0] *= F(1.6)
    _check_backward(bt, synth.make_patch(40, seed=71), 0.4, 71)


def test_forward_backward_more_than_1024_boxes():
    # the work-item prefix tables no longer fit shared memory (kMaxBaseSmem): global-memory search paths
    bt = synth.make_batch(160, 64, 64, seed=91, max_boxes=8, min_boxes=7)
    assert bt.boxes.shape[0] > 1024
    _check_backward(bt, synth.make_patch(20, seed=91), 0.5, 91)
