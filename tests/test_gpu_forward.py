"""GPU parity: CUDA forward (through the C ABI) vs the oracle.  Bars (BASELINE.json north_star):
bit-exact box/transform indexing and masks; warped/composited images within 1e-5 max-abs (float32).
The kernels are built to reproduce the oracle's float32 op order, so equality is asserted exactly
and the 1e-5 bar is reported as head-room."""
import os

import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import ops, synth
from oracle import patcher
from tests._util import run_forward, to_device

pytestmark = pytest.mark.gpu
F = np.float32
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_geometry(bt, scale, H, W, tol):
    rows = []
    for b in range(len(bt.offsets) - 1):
        for box, p in zip(bt.boxes_of(b), bt.params_of(b)):
            sc = p["scale"] if p["scale"] >= 0 else scale
            try:
                pl = patcher.create(box, sc, p["uy"], p["ux"], tol, H, W)
                rows.append([pl.y0, pl.x0, pl.ps, pl.d, pl.pad_lo, pl.pad_hi, int(pl.valid)])
            except ValueError:
                rows.append(None)
    return rows


@pytest.mark.parametrize("H,scale,seed", [(512, 0.4, 1), (640, 0.4, 2), (1024, 0.25, 3), (96, 0.9, 4), (64, 0.05, 5)])
def test_box_geometry_bit_exact(H, scale, seed):
    bt = synth.make_batch(16, H, H, seed=seed, max_boxes=8)
    # add degenerate / border boxes to the first image's neighbours
    bt.boxes[0] = [0, 0, 5, 5]
    bt.boxes[1] = [H - 9, H - 9, H, H]
    d = to_device(bt)
    sc = torch.tensor(scale, dtype=torch.float32, device="cuda")
    g = ops.box_geometry(bt.images.shape, 100, d["boxes"], d["offsets"], d["params"], sc).cpu().numpy()
    ref = _oracle_geometry(bt, scale, H, H, 0.2)
    assert len(ref) == len(g)
    for r, row in zip(ref, g):
        if r is None:
            assert row[6] == 0
        else:
            assert list(row[:7]) == r


def _check_forward(bt, patch, scale, geom=None, want_mask=False, per_image=False):
    geom = geom or ops.PatchGeometry()
    out, mask, ctx, d = run_forward(patch, scale, bt, geom, want_mask)
    ops.check_workspace(ctx)
    bx, pr = bt.ragged()
    ref, ref_mask, _ = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, scale, tolerance=geom.tolerance,
                                               noise_amp=geom.noise_amp, min_patch_area=geom.min_patch_area,
                                               want_mask=want_mask)
    got = out.cpu().numpy()
    diff = np.abs(got - ref)
    assert diff.max() <= 1e-5, f"max-abs {diff.max()}"
    nneq = int((got != ref).sum())
    assert nneq == 0, f"{nneq} of {got.size} elements differ in the last bits (max-abs {diff.max()})"
    if want_mask:
        np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask)
    return got


def test_forward_matches_committed_golden_vector():
    g = np.load(os.path.join(GOLD, "oracle_small.npz"))
    bt = synth.make_batch(2, 64, 64, max_boxes=3, min_boxes=2, seed=314)
    got = _check_forward(bt, synth.make_patch(16, seed=3), 0.4)
    np.testing.assert_array_equal(got, g["out"])


@pytest.mark.parametrize("B,H,P,scale,seed", [
    (4, 256, 100, 0.4, 21),      # upsampling resize (ps > P for tall boxes)
    (3, 320, 300, 0.4, 22),      # downsampling, antialias spans > 3
    (2, 200, 640, 0.3, 23),      # script default texture, strong downsampling
    (5, 128, 24, 0.9, 24),       # large scale: diag clamps to W
    (3, 250, 50, 0.4, 25),       # H*W not a multiple of 4 -> scalar image pass
])
def test_forward_bit_exact(B, H, P, scale, seed):
    bt = synth.make_batch(B, H, H, seed=seed, max_boxes=6)
    _check_forward(bt, synth.make_patch(P, seed=seed), scale)


def test_forward_perspective_row():
    bt = synth.make_batch(3, 256, 256, seed=31, max_boxes=4, perspective=2e-4)
    _check_forward(bt, synth.make_patch(100, seed=31), 0.4)


def test_forward_zero_boxes_is_identity_and_empty_rows():
    bt = synth.make_batch(3, 128, 128, seed=41, max_boxes=0)
    got = _check_forward(bt, synth.make_patch(32), 0.4)
    np.testing.assert_array_equal(got, bt.images)
    # ragged: only the middle image has boxes
    bt2 = synth.make_batch(3, 128, 128, seed=42, max_boxes=3, min_boxes=3)
    keep = slice(bt2.offsets[1], bt2.offsets[2])
    bt2.boxes, bt2.params = bt2.boxes[keep], bt2.params[keep]
    n = bt2.boxes.shape[0]
    bt2.offsets = np.array([0, 0, n, n], dtype=np.int32)
    got = _check_forward(bt2, synth.make_patch(32), 0.4)
    np.testing.assert_array_equal(got[0], bt2.images[0])
    np.testing.assert_array_equal(got[2], bt2.images[2])


def test_forward_in_place_output_aliases_input():
    bt = synth.make_batch(2, 128, 128, seed=43, max_boxes=3)
    patch = synth.make_patch(32)
    ref, _, _, _ = run_forward(patch, 0.4, bt)
    d = to_device(bt)
    img = d["images"].clone()
    sc = torch.tensor(0.4, dtype=torch.float32, device="cuda")
    out, _, _ = ops.apply_forward(torch.from_numpy(patch).cuda(), sc, img, d["boxes"], d["offsets"], d["params"],
                                  d["print_wb"], out=img)
    torch.cuda.synchronize()
    assert out.data_ptr() == img.data_ptr()
    assert torch.equal(out, ref)


def test_masker_training_variant_strided_flipped_patches_and_mask():
    # attack_detection.py:478-498: patches = shuffled 240x240 crops of the batch, randomly flipped
    B, H, P = 4, 320, 240
    bt = synth.make_batch(B, H, H, seed=51, max_boxes=3, scale_range=(0.3, 0.5))
    geom = ops.PatchGeometry(tolerance=0.5, noise_amp=0.1, max_scale=0.5)
    imgs = torch.from_numpy(bt.images).cuda()
    perm = [2, 0, 3, 1]
    crops = imgs[perm][:, :P, :P, :]                    # gathered copy (shuffle), then zero-copy flips
    view = crops.flip(1)                                # torch.flip copies; emulate the view with numpy instead
    patches_np = bt.images[perm][:, :P, :P, :][:, ::-1]
    # non-contiguous, negative row stride is not expressible in torch: use a contiguous tensor here and a
    # strided (sliced) one below
    _check_forward(bt, np.ascontiguousarray(patches_np), 0.4, geom, want_mask=True)
    assert torch.equal(view.cpu(), torch.from_numpy(np.ascontiguousarray(patches_np)))
    # strided view without a copy: crops of the images tensor itself (row stride W*3, batch stride H*W*3)
    strided = imgs[:, :P, :P, :]
    out, mask, ctx, _ = run_forward(None, 0.4, bt, geom, want_mask=True, patch_t=strided)
    bx, pr = bt.ragged()
    ref, ref_mask, _ = patcher.patcher_forward(bt.images[:, :P, :P, :], bt.images, bx, pr, bt.print_wb, 0.4,
                                               tolerance=0.5, noise_amp=0.1, want_mask=True)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask)


def test_forward_config1_shape():
    # BASELINE.json configs[0]: batch 8, 512x512, 100x100 patch
    bt = synth.make_batch(8, 512, 512, seed=1234, max_boxes=8)
    _check_forward(bt, synth.make_patch(100), 0.4)


def test_cpu_tensors_are_rejected_loudly():
    bt = synth.make_batch(1, 64, 64, seed=1, max_boxes=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        ops.apply_forward(torch.zeros(8, 8, 3), torch.tensor(0.4), torch.from_numpy(bt.images),
                          torch.from_numpy(bt.boxes), torch.from_numpy(bt.offsets),
                          torch.zeros(1, 48, dtype=torch.uint8), torch.from_numpy(bt.print_wb))


def test_forward_out_of_range_image_clips_whole_window():
    # attacker.py:441 clips the WHOLE d x d window, background included; pixels outside every window keep values > 1
    bt = synth.make_batch(3, 160, 160, seed=71, max_boxes=3, min_boxes=1)
    bt.images[0] *= F(1.6)
    bt.images[2, :40] = F(-1.3)
    got = _check_forward(bt, synth.make_patch(40, seed=71), 0.4)
    assert np.abs(got[0]).max() > 1.0          # untouched background outside the windows
    assert np.abs(got[1]).max() <= 1.0


def test_forward_many_mutually_overlapping_boxes():
    # 14 boxes crowded into one image: every pixel sees several windows; the last paste must win per channel
    bt = synth.make_batch(2, 192, 192, seed=72, max_boxes=14, min_boxes=14)
    rng = np.random.default_rng(72)
    for j in range(bt.offsets[0], bt.offsets[1]):
        cy, cx = rng.uniform(70, 120, 2)
        h, w = rng.uniform(60, 130), rng.uniform(30, 70)
        bt.boxes[j] = [max(cy - h / 2, 0), max(cx - w / 2, 0), min(cy + h / 2, 192), min(cx + w / 2, 192)]
    _check_forward(bt, synth.make_patch(36, seed=72), 0.4)


def test_masker_mask_with_overlapping_windows():
    # mask = original - pasted over the union of the windows (attack_detection.py:429-430), later pastes on top
    bt = synth.make_batch(2, 160, 160, seed=73, max_boxes=3, min_boxes=3)
    for b in range(2):
        bt.boxes[bt.offsets[b] + 1] = bt.boxes[bt.offsets[b]] + F(6.0)
    bt.boxes = np.clip(bt.boxes, 0, 160).astype(F)
    _check_forward(bt, synth.make_patch(32, seed=73), 0.4, geom=ops.PatchGeometry(tolerance=0.5, noise_amp=0.1), want_mask=True)


def test_cuda_forward_equals_reference_patcher_fixture():
    """CUDA path vs the output of the reference's own Patcher.call run on the NumPy TF shim (tests/golden/patcher_ref.npz)."""
    from tests.test_reference_patcher_fixture import load_patcher_fixture
    g, _, _ = load_patcher_fixture()
    dev = "cuda"
    params = np.ascontiguousarray(g["params"]).view(patcher.BOX_PARAMS).reshape(-1)
    out, _, ctx = ops.apply_forward(torch.from_numpy(g["patch"]).to(dev), torch.tensor(float(g["scale"]), device=dev),
                                    torch.from_numpy(g["images"]).to(dev), torch.from_numpy(g["boxes"]).to(dev),
                                    torch.from_numpy(g["offsets"]).to(dev), ops.params_to_tensor(params, dev),
                                    torch.from_numpy(g["print_wb"]).to(dev), ops.PatchGeometry())
    ops.check_workspace(ctx)
    np.testing.assert_array_equal(out.cpu().numpy(), g["out_ref"])
    b = np.load(os.path.join(GOLD, "brightness_ref.npz"))
    got = ops.brightness_match(torch.from_numpy(b["src"]).to(dev), torch.from_numpy(b["tgt"]).to(dev)).cpu().numpy()
    np.testing.assert_array_equal(got, b["out_ref"])


@pytest.mark.parametrize("tag,tol", [("eval", 0.0), ("train", 0.5)])
def test_cuda_masker_equals_reference_masker_fixture(tag, tol):
    """CUDA Masker path (mask output, per-image patch textures) vs the reference's own Masker.call on the NumPy TF shim."""
    from tests.test_reference_patcher_fixture import load_masker_fixture
    g, _, _, params = load_masker_fixture(tag)
    dev = "cuda"
    geom = ops.PatchGeometry(tolerance=tol, noise_amp=0.1, max_scale=0.5 if tag == "train" else 1.0)
    out, mask, ctx = ops.apply_forward(torch.from_numpy(g[f"{tag}_patch"]).to(dev), torch.tensor(float(g[f"{tag}_scale"]), device=dev),
                                       torch.from_numpy(g[f"{tag}_images"]).to(dev), torch.from_numpy(g[f"{tag}_boxes"]).to(dev),
                                       torch.from_numpy(g[f"{tag}_offsets"]).to(dev), ops.params_to_tensor(params, dev),
                                       torch.from_numpy(g[f"{tag}_print_wb"]).to(dev), geom, want_mask=True)
    ops.check_workspace(ctx)
    np.testing.assert_array_equal(out.cpu().numpy(), g[f"{tag}_out"])
    np.testing.assert_array_equal(mask.cpu().numpy(), g[f"{tag}_mask"])


def test_cuda_forward_equals_reference_patcher_fixture_harder_case():
    """patcher_ref2 (tests/golden): overlapping boxes, border-clamped windows, down- and up-sampling, a filtered box --
    CUDA path vs the reference's own Patcher.call run on the NumPy TF shim."""
    g = np.load(os.path.join(GOLD, "patcher_ref2.npz"))
    params = np.ascontiguousarray(g["params"]).view(patcher.BOX_PARAMS).reshape(-1)
    dev = "cuda"
    out, _, ctx = ops.apply_forward(torch.from_numpy(g["patch"]).to(dev), torch.tensor(float(g["scale"]), device=dev),
                                    torch.from_numpy(g["images"]).to(dev), torch.from_numpy(g["boxes"]).to(dev),
                                    torch.from_numpy(g["offsets"]).to(dev), ops.params_to_tensor(params, dev),
                                    torch.from_numpy(g["print_wb"]).to(dev), ops.PatchGeometry())
    ops.check_workspace(ctx)
    np.testing.assert_array_equal(out.cpu().numpy(), g["out_ref"])
