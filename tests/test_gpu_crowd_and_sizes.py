"""GPU parity at the edges the reference reaches but small synthetic batches do not:

* crowds: 40 and 100 person boxes on ONE image (100 = nms_configs.max_output_size,
  automl/efficientdet/hparams_config.py:265 -- the most boxes `Patcher` can ever be handed per image), all
  overlapping: forward bit-exact, backward within 1e-4 relative L2, nothing dropped;
* the shapes of BASELINE.json configs[2] (perspective EOT at 512x512, P=100) and configs[3] (1024x1024, P=300,
  8 boxes per image), forward + backward against the oracle;
* the serial resize adjoint at the scripts' default patch texture (P=640);
* total_boxes as a capacity (the count in use is read on the device) and the device-side transform draw.
"""
import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import ops, synth
from mladversarialobjectdetection_b200.sampler import TransformSampler
from oracle import patcher
from tests._util import run_forward, to_device
from tests.test_gpu_backward import _check_backward

pytestmark = pytest.mark.gpu
F = np.float32


def _crowd(n_boxes, H, seed):
    """one image with n_boxes mutually overlapping person boxes + one ordinary image"""
    bt = synth.make_batch(2, H, H, seed=seed, max_boxes=n_boxes, min_boxes=n_boxes)
    rng = np.random.default_rng(seed)
    for j in range(bt.offsets[0], bt.offsets[1]):
        cy, cx = rng.uniform(0.3 * H, 0.7 * H, 2)
        h, w = rng.uniform(0.15 * H, 0.5 * H), rng.uniform(0.08 * H, 0.25 * H)
        bt.boxes[j] = [max(cy - h / 2, 0), max(cx - w / 2, 0), min(cy + h / 2, H), min(cx + w / 2, H)]
    return bt


@pytest.mark.parametrize("n_boxes", [33, 40, 100])
def test_crowd_on_one_image_forward_bit_exact_backward_rel_l2(n_boxes):
    bt = _crowd(n_boxes, 512, 100 + n_boxes)
    patch = synth.make_patch(48, seed=n_boxes)
    gp, ctx, _, _ = _check_backward(bt, patch, 0.4, n_boxes)        # asserts forward == oracle bit for bit, too
    ops.check_workspace(ctx)                                        # nothing flagged, nothing dropped


def test_crowd_with_out_of_range_image():
    bt = _crowd(40, 256, 140)
    bt.images[0] *= F(1.5)
    _check_backward(bt, synth.make_patch(32, seed=5), 0.4, 141)


def test_config3_shape_perspective_512_p100():
    bt = synth.make_batch(3, 512, 512, seed=301, max_boxes=8, min_boxes=3, perspective=2e-4)
    _check_backward(bt, synth.make_patch(100, seed=301), 0.4, 301, smooth_grad=True)


def test_config4_shape_1024_p300_8_boxes():
    bt = synth.make_batch(2, 1024, 1024, seed=401, max_boxes=8, min_boxes=8)
    _check_backward(bt, synth.make_patch(300, seed=401), 0.4, 401)


def test_serial_adjoint_at_p640():
    bt = synth.make_batch(2, 320, 320, seed=641, max_boxes=5, min_boxes=4)
    patch = synth.make_patch(640, seed=641)
    g1, _, _, _ = _check_backward(bt, patch, 0.4, 641)
    g2, _, _, _ = _check_backward(bt, patch, 0.4, 641, geom=ops.PatchGeometry(serial_adjoint=True))
    assert torch.allclose(g1, g2, rtol=1e-5, atol=1e-7 * float(g1.abs().max()))


def test_total_boxes_is_a_capacity():
    """boxes / params padded past row_splits[-1] (what a sync-free first pass hands over): same images, same gradient"""
    bt = synth.make_batch(3, 160, 160, seed=77, max_boxes=4, min_boxes=2)
    patch = synth.make_patch(40, seed=77)
    out, _, ctx, d = run_forward(patch, 0.4, bt)
    n, cap = bt.boxes.shape[0], bt.boxes.shape[0] + 7
    boxes = torch.full((cap, 4), 123.0, device="cuda")
    boxes[:n] = d["boxes"]
    params = torch.randint(0, 255, (cap, 48), dtype=torch.uint8, device="cuda")
    params[:n] = d["params"]
    sc = torch.tensor(0.4, device="cuda")
    out2, _, ctx2 = ops.apply_forward(torch.from_numpy(patch).cuda(), sc, d["images"], boxes, d["offsets"], params, d["print_wb"])
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    ops.check_workspace(ctx2)
    G = torch.randn_like(out)
    assert torch.equal(ops.apply_backward(ctx, G), ops.apply_backward(ctx2, G))
    # more boxes than the capacity: flagged on the checking call, never silent
    _, _, ctx3 = ops.apply_forward(torch.from_numpy(patch).cuda(), sc, d["images"], d["boxes"][:n - 1], d["offsets"],
                                   d["params"][:n - 1], d["print_wb"])
    with pytest.raises(RuntimeError, match="capacity"):
        ops.check_workspace(ctx3)


def test_draw_kernel_equals_the_torch_definition():
    """eot_draw_transforms (one launch) against sampler.box_params / print_wb (the same hash in torch ops)"""
    bt = synth.make_batch(7, 64, 64, seed=9, max_boxes=5)
    off = torch.from_numpy(bt.offsets).cuda()
    n = int(bt.offsets[-1])
    for kw, rng_ in [(dict(), None), (dict(perspective=3e-4), None), (dict(), (0.3, 0.5))]:
        s = TransformSampler(seed=12, **kw)
        params, wb = s.draw(5, 40, off, n + 3, scale_range=rng_)
        ref = s.box_params(5, 40, off, n, scale_range=rng_).view(torch.float32).reshape(n, 12)
        got = params.view(torch.float32).reshape(n + 3, 12)
        torch.cuda.synchronize()
        assert float(got[n:].abs().max()) == 0.0                     # unused slots are zero
        exact = [0, 1, 2, 5, 6, 7]                                   # jitter, delta, projective row, scale: same float32 ops
        assert torch.equal(got[:n, exact], ref[:, exact])
        assert torch.equal(got[:n, 8:10].view(torch.int32), ref[:, 8:10].view(torch.int32))   # Philox keys
        assert torch.allclose(got[:n, 3:5], ref[:, 3:5], rtol=0, atol=2e-7)                      # cos / sin (libdevice)
        assert torch.allclose(wb, s.print_wb(5, 40, 7, "cuda"), rtol=0, atol=2e-7)


def test_no_writes_outside_the_callers_buffers():
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are hunted with red zones of our own: every
    buffer the C ABI writes (workspace, out images, masks, gradient) sits between canary blocks inside a larger
    allocation, at odd sizes, and the canaries must survive a forward + backward (crowd, projective rows, out-of-range
    image, mask output)."""
    import ctypes
    from mladversarialobjectdetection_b200 import _lib

    def guarded(nbytes, align=256, pad=4096):
        raw = torch.full((nbytes + 2 * pad + align,), 0xA5, dtype=torch.uint8, device="cuda")
        off = pad + (-(raw.data_ptr() + pad)) % align
        return raw, off, raw[off:off + nbytes]

    def canaries_intact(raw, off, nbytes):
        return bool((raw[:off] == 0xA5).all()) and bool((raw[off + nbytes:] == 0xA5).all())

    for want_mask, persp, H, P, nb in [(False, 0.0, 150, 37, 9), (True, 3e-4, 131, 20, 5), (False, 2e-4, 97, 64, 40)]:
        bt = synth.make_batch(3, H, H, seed=H, max_boxes=nb, min_boxes=max(1, nb // 2), perspective=persp)
        bt.images[1] *= F(1.4)
        d = to_device(bt)
        patch = torch.from_numpy(synth.make_patch(P, seed=1)).cuda()
        sc = torch.tensor(0.45, device="cuda")
        geom = ops.PatchGeometry()
        shape = ops._shape(d["images"], patch, int(bt.boxes.shape[0]), geom, want_mask)
        need = ops.workspace_bytes(shape)
        ws_raw, ws_off, ws = guarded(need)
        nimg = d["images"].numel() * 4
        out_raw, out_off, out_b = guarded(nimg, align=16)
        out = out_b.view(torch.float32).view_as(d["images"])
        lib = _lib.load()
        mask_raw = mask_off = None
        mask = None
        if want_mask:
            mask_raw, mask_off, mask_b = guarded(nimg, align=16)
            mask = mask_b.view(torch.float32).view_as(d["images"])
        _lib.check(lib.eot_apply_fwd(ctypes.byref(shape), ops._ptr(patch), ops._ptr(sc), ops._ptr(d["images"]), ops._ptr(d["boxes"]),
                                     ops._ptr(d["offsets"]), ops._ptr(d["params"]), ops._ptr(d["print_wb"]), ops._ptr(out),
                                     ops._ptr(mask), ops._ptr(ws), ctypes.c_size_t(need), ops._stream()), "eot_apply_fwd")
        gp_raw, gp_off, gp_b = guarded(P * P * 3 * 4, align=16)
        if not want_mask:
            G = torch.randn_like(out)
            _lib.check(lib.eot_apply_bwd(ctypes.byref(shape), ops._ptr(patch), ops._ptr(d["print_wb"]), ops._ptr(G), ops._ptr(ws),
                                         ctypes.c_size_t(need), ops._ptr(gp_b), 0, ops._stream()), "eot_apply_bwd")
        torch.cuda.synchronize()
        assert canaries_intact(ws_raw, ws_off, need), "write outside the workspace"
        assert canaries_intact(out_raw, out_off, nimg), "write outside out_images"
        assert canaries_intact(gp_raw, gp_off, P * P * 3 * 4), "write outside grad_patch"
        if want_mask:
            assert canaries_intact(mask_raw, mask_off, nimg), "write outside out_masks"
        # and the guarded run computes what the ordinary call computes
        ref_out, ref_mask, _ = ops.apply_forward(patch, sc, d["images"], d["boxes"], d["offsets"], d["params"], d["print_wb"], geom,
                                                 want_mask=want_mask)
        assert torch.equal(out, ref_out)
        if want_mask:
            assert torch.equal(mask, ref_mask)
