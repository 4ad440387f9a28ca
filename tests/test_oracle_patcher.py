"""Oracle self-checks for Patcher / BrightnessMatcher restatement (no reference test pins these:
SURVEY.md section 4 -- analytic identities, hand-computed geometry, autograd cross-check)."""
import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import synth
from oracle import patcher, tfops

F = np.float32


def _params(n=1, **kw):
    p = np.zeros(n, dtype=patcher.BOX_PARAMS)
    p["uy"] = 0.5; p["ux"] = 0.5; p["cos"] = 1.0; p["sin"] = 0.0; p["scale"] = -1.0
    for k, v in kw.items():
        p[k] = v
    return p


def test_box_params_dtype_matches_product_record():
    assert patcher.BOX_PARAMS == synth.BOX_PARAMS and patcher.BOX_PARAMS.itemsize == 48


def test_geometry_worked_example():
    # 640^2 image, box (50,125,400,200), scale .4, zero jitter
    pl = patcher.create([50, 125, 400, 200], 0.4, 0.5, 0.5, 0.2, 640, 640)
    assert (pl.y0, pl.x0, pl.ps, pl.d) == (126, 63, 140, 197)
    assert (pl.pad_lo, pl.pad_hi) == (28, 29) and pl.valid
    np.testing.assert_allclose(pl.row, [126.005, 63.505, 140, 140, 197.98988], rtol=1e-6)


def test_geometry_clamps_to_image_and_filters_small():
    pl = patcher.create([400, 500, 640, 640], 0.4, 0.999, 0.999, 0.2, 640, 640)
    assert pl.y0 + pl.d <= 640 and pl.x0 + pl.d <= 640
    pl = patcher.create([0, 0, 5, 5], 0.4, 0.5, 0.5, 0.2, 640, 640)       # ps = 2 -> area 4, not > 4
    assert pl.ps == 2 and not pl.valid
    pl = patcher.create([0, 0, 8, 8], 0.4, 0.0, 0.0, 0.2, 640, 640)       # ps = 3 -> kept
    assert pl.valid and (pl.y0, pl.x0, pl.d, pl.pad_lo, pl.pad_hi) == (1, 1, 4, 0, 1)
    pl = patcher.create([0, 0, 8, 2], 0.4, 0.0, 0.0, 0.2, 640, 640)       # x centre 1 - .2 -> clamped at 0
    assert pl.valid and pl.x0 == 0


def test_zero_boxes_is_identity():
    bt = synth.make_batch(2, 64, 64, max_boxes=0)
    bx, pr = bt.ragged()
    out, _, _ = patcher.patcher_forward(synth.make_patch(16), bt.images, bx, pr, bt.print_wb, 0.4)
    np.testing.assert_array_equal(out, bt.images)


def test_identity_transform_pastes_clipped_resized_patch():
    H = W = 96
    img = np.random.default_rng(0).uniform(-1, 1, (H, W, 3)).astype(F)
    patch = synth.make_patch(20)
    wb = np.array([1, 1, 1, 0, 0, 0], dtype=F)
    box = np.array([[10, 20, 90, 60]], dtype=F)
    out, _, st = patcher.add_patches_to_image(img, patch, box, _params(), wb, 0.4, noise_amp=0.0)
    pl = st.boxes[0].plan
    m = patcher.brightness_match(patch, img)
    r = np.clip(tfops.aa_resize(m, pl.ps, pl.ps), -1, 1)
    y, x = pl.y0 + pl.pad_lo, pl.x0 + pl.pad_lo
    np.testing.assert_array_equal(out[y:y + pl.ps, x:x + pl.ps], r)
    # outside the ps x ps core nothing changed (pad = -2 -> background)
    keep = np.ones((H, W), bool); keep[y:y + pl.ps, x:x + pl.ps] = False
    np.testing.assert_array_equal(out[keep], img[keep])


def test_sequential_paste_later_box_wins():
    H = W = 128
    img = np.zeros((H, W, 3), F)
    patch = np.ones((8, 8, 3), F)
    wb = np.array([1, 1, 1, 0, 0, 0], dtype=F)
    boxes = np.array([[20, 20, 100, 100], [30, 30, 110, 110]], dtype=F)
    p = _params(2, delta=[-0.5, 0.25])
    out, _, st = patcher.add_patches_to_image(img, patch, boxes, p, wb, 0.4, noise_amp=0.0)
    a, b = st.boxes[0].plan, st.boxes[1].plan
    ya, yb = a.y0 + a.pad_lo, b.y0 + b.pad_lo
    inter_y = max(ya, yb) + 1
    # in the overlap of the two cores the second box's value is visible
    assert out[inter_y, inter_y, 0] == st.boxes[1].R[inter_y - b.y0, inter_y - b.x0, 0]
    assert out[ya + 1, ya + 1, 0] == st.boxes[0].R[ya + 1 - a.y0, ya + 1 - a.x0, 0]
    assert st.boxes[0].R[ya + 1 - a.y0, ya + 1 - a.x0, 0] != st.boxes[1].R[inter_y - b.y0, inter_y - b.x0, 0]


def test_brightness_match_shifts_mean_luma():
    rng = np.random.default_rng(4)
    src = rng.uniform(-0.3, 0.3, (32, 32, 3)).astype(F)
    tgt = rng.uniform(0.2, 0.6, (40, 40, 3)).astype(F)
    out = patcher.brightness_match(src, tgt)
    def luma(x):
        return float(tfops.dot3((x + 1) * tfops.C_127_255, tfops.RGB2YUV)[..., 0].mean())
    assert abs(luma(out) - luma(tgt)) < 2e-3


def test_masker_mask_is_original_minus_pasted_inside_windows():
    bt = synth.make_batch(2, 96, 96, max_boxes=3, seed=5, scale_range=(0.3, 0.5))
    bx, pr = bt.ragged()
    patches = np.stack([bt.images[1, :24, :24], bt.images[0, :24, :24]])
    out, mask, states = patcher.patcher_forward(patches, bt.images, bx, pr, bt.print_wb, 0.4,
                                                tolerance=0.5, noise_amp=0.1, want_mask=True)
    for b in range(2):
        inwin = np.zeros((96, 96), bool)
        for bs in states[b].boxes:
            inwin[bs.plan.y0:bs.plan.y0 + bs.plan.d, bs.plan.x0:bs.plan.x0 + bs.plan.d] = True
        np.testing.assert_array_equal(mask[b][~inwin], 0)
        np.testing.assert_array_equal(mask[b][inwin], (bt.images[b] - out[b])[inwin])


# ---- backward: oracle's hand-written chain vs torch autograd (float64) of the same forward, with
# ---- the rotate op carrying TF's REGISTERED gradient (inverse warp of the gradient, fill 0)
class _TFRotate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pad, T):
        ctx.T = T
        return torch.from_numpy(tfops.projective_bilinear(pad.numpy().astype(F), T, -2.0).astype(np.float64))

    @staticmethod
    def backward(ctx, g):
        return torch.from_numpy(tfops.projective_bilinear_grad(g.numpy().astype(F), ctx.T).astype(np.float64)), None


def _dense_resize_matrix(out, inp):
    s, w, n = tfops.compute_spans(out, inp)
    M = np.zeros((out, inp))
    for o in range(out):
        for k in range(n):
            if s[o] + k < inp:
                M[o, s[o] + k] += w[o, k]
    return torch.from_numpy(M)


def _torch_forward_image(patch_t, image, boxes, params, wb, scale, states, noise_amp):
    K = torch.from_numpy(tfops.RGB2YUV.astype(np.float64))
    Ki = torch.from_numpy(tfops.YUV2RGB.astype(np.float64))
    img = torch.from_numpy(image.astype(np.float64))
    w = torch.from_numpy(wb[:3].astype(np.float64)); b = torch.from_numpy(wb[3:].astype(np.float64))
    q = torch.clamp(w * patch_t + b, -1, 1)
    s = (q + 1) * float(tfops.C_127_255)
    yuv = s @ K
    t = (img + 1) * float(tfops.C_127_255)
    mu_t = (t @ K)[..., 0].mean()
    yp = torch.clamp(yuv[..., 0] - yuv[..., 0].mean() + mu_t, 0, 1)
    rgb = torch.stack([yp, yuv[..., 1], yuv[..., 2]], -1) @ Ki
    m = torch.clamp(rgb, 0, 1) * float(tfops.C_255_127) - 1
    out = img.clone()
    P = patch_t.shape[0]
    for bs, prm in zip(states.boxes, [p for p in params]):
        pl = bs.plan
        Wm = _dense_resize_matrix(pl.ps, P)
        r = torch.einsum("oi,ijc->ojc", Wm, m)
        r = torch.einsum("pj,ojc->opc", Wm, r)
        noise = tfops.uniform_noise(pl.ps * pl.ps * 3, int(prm["key0"]), int(prm["key1"]), noise_amp)
        u = torch.clamp(r + torch.from_numpy(noise.reshape(pl.ps, pl.ps, 3).astype(np.float64)) + float(prm["delta"]), -1, 1)
        pad = torch.nn.functional.pad(u, (0, 0, pl.pad_lo, pl.pad_hi, pl.pad_lo, pl.pad_hi), value=-2.0)
        R = _TFRotate.apply(pad, bs.T)
        ys, xs = slice(pl.y0, pl.y0 + pl.d), slice(pl.x0, pl.x0 + pl.d)
        bg = out[ys, xs]
        o = torch.clamp(torch.where(R < -1, bg, R), -1, 1)
        out = out.clone()
        out[ys, xs] = o
    return out


@pytest.mark.parametrize("seed,P,H", [(11, 24, 96), (12, 40, 128)])
def test_backward_matches_autograd_with_tf_rotate_gradient(seed, P, H):
    bt = synth.make_batch(2, H, H, max_boxes=3, min_boxes=2, seed=seed, image_fill="smooth")
    bx, pr = bt.ragged()
    patch = synth.make_patch(P, seed=seed)
    out, _, states = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, 0.4)
    G = np.random.default_rng(seed).normal(size=out.shape).astype(F)
    g_or = patcher.patcher_backward(G, patch, bt.print_wb, states, dtype=np.float64)
    pt = torch.tensor(patch.astype(np.float64), requires_grad=True)
    tot = 0
    for b in range(2):
        valid_params = [p for p, box in zip(pr[b], bx[b])
                        if patcher.create(box, 0.4, p["uy"], p["ux"], 0.2, H, H).valid]
        o = _torch_forward_image(pt, bt.images[b], bx[b], valid_params, bt.print_wb[b], 0.4, states[b], 0.01)
        assert np.abs(o.detach().numpy() - out[b]).max() < 1e-4        # same forward
        tot = tot + (o * torch.from_numpy(G[b].astype(np.float64))).sum()
    tot.backward()
    g_ag = pt.grad.numpy()
    rel = np.linalg.norm(g_or - g_ag) / np.linalg.norm(g_ag)
    assert rel < 1e-4, rel
    # float32 backward agrees with float64 backward
    g32 = patcher.patcher_backward(G, patch, bt.print_wb, states, dtype=F)
    assert np.linalg.norm(g32 - g_or) / np.linalg.norm(g_or) < 1e-5
