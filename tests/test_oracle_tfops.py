"""Oracle self-checks: restated TF/TFA ops vs independent implementations available here."""
import numpy as np
import pytest
import torch

from oracle import tfops

F = np.float32


@pytest.mark.parametrize("P,ps", [(640, 140), (100, 37), (100, 250), (300, 120), (100, 100), (64, 3)])
def test_aa_resize_matches_torch_antialias(P, ps):
    # formula == torch bilinear antialias to <= 5e-7
    x = np.random.default_rng(P + ps).uniform(-1, 1, (P, P, 3)).astype(F)
    y = tfops.aa_resize(x, ps, ps)
    yt = torch.nn.functional.interpolate(torch.from_numpy(x).permute(2, 0, 1)[None], size=(ps, ps),
                                         mode="bilinear", antialias=True, align_corners=False)
    yt = yt[0].permute(1, 2, 0).numpy()
    assert np.abs(y - yt).max() <= 5e-7


def test_span_weights_normalised_and_span_size():
    starts, w, n = tfops.compute_spans(32, 640)       # 20x down: span 2*20+1
    assert n == 41
    np.testing.assert_allclose(w.sum(1), 1.0, atol=1e-6)
    starts, w, n = tfops.compute_spans(250, 100)      # upsample: plain bilinear, span 3
    assert n == 3
    assert starts.min() == 0 and (starts + n).max() <= 100 + n


@pytest.mark.parametrize("P,ps", [(100, 37), (50, 120)])
def test_aa_resize_grad_is_exact_transpose(P, ps):
    rng = np.random.default_rng(1)
    x = rng.normal(size=(P, P, 3))
    g = rng.normal(size=(ps, ps, 3))
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    yt = torch.nn.functional.interpolate(xt.permute(2, 0, 1)[None], size=(ps, ps), mode="bilinear",
                                         antialias=True, align_corners=False)[0].permute(1, 2, 0)
    yt.backward(torch.tensor(g))
    gx = tfops.aa_resize_grad(g, P, P, dtype=np.float64)
    rel = np.linalg.norm(gx - xt.grad.numpy()) / np.linalg.norm(gx)
    assert rel < 1e-5      # float32 span weights (TF) vs float64 torch weights


def test_projective_identity_and_rotation_zero():
    img = np.random.default_rng(2).uniform(-1, 1, (31, 31, 3)).astype(F)
    T = tfops.rotation_transform(F(1), F(0), 31)
    np.testing.assert_array_equal(T, np.array([1, -0.0, 0, 0, 1, 0, 0, 0], dtype=F))
    np.testing.assert_array_equal(tfops.projective_bilinear(img, T, -2.0), img)


def test_projective_matches_grid_sample_loosely():
    # grid_sample(align_corners=True) on (x+2) agrees to ~2e-5
    D = 97
    img = np.random.default_rng(3).uniform(-1, 1, (D, D, 3)).astype(F)
    th = 15 * np.pi / 180
    T = tfops.rotation_transform(F(np.cos(th)), F(np.sin(th)), D)
    R = tfops.projective_bilinear(img, T, -2.0)
    ix, iy, _ = tfops.projective_coords(T, D, D)
    grid = np.stack([ix / (D - 1) * 2 - 1, iy / (D - 1) * 2 - 1], -1)[None]
    t = torch.from_numpy(img + 2).permute(2, 0, 1)[None].double()
    Rt = torch.nn.functional.grid_sample(t, torch.from_numpy(grid).double(), mode="bilinear",
                                         padding_mode="zeros", align_corners=True)[0].permute(1, 2, 0).numpy() - 2
    assert np.abs(R - Rt).max() < 2e-4
    assert 0.05 < (R < -1).mean() < 0.2


def test_invert_transform_roundtrip():
    T = tfops.rotation_transform(F(np.cos(0.3)), F(np.sin(0.3)), 197, F(2e-4), F(-1e-4))
    Ti = tfops.invert_transform(T)
    M = np.append(T, 1).reshape(3, 3).astype(np.float64)
    Mi = np.append(Ti, 1).reshape(3, 3).astype(np.float64)
    prod = M @ Mi
    np.testing.assert_allclose(prod / prod[2, 2], np.eye(3), atol=2e-5)


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32 10, counter 0, key 0
    out = tfops.philox4x32_10(np.array([0], dtype=np.uint32), 0, 0)[0]
    assert [int(v) for v in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_uniform_noise_range_and_determinism():
    n = tfops.uniform_noise(3001, 123, 456, 0.01)
    assert n.dtype == F and n.shape == (3001,)
    assert n.min() >= -0.01 and n.max() < 0.01
    assert abs(float(n.mean())) < 1e-3
    np.testing.assert_array_equal(n, tfops.uniform_noise(3001, 123, 456, 0.01))
    assert not np.array_equal(n, tfops.uniform_noise(3001, 124, 456, 0.01))


def test_total_variation_gradient_numeric():
    x = np.random.default_rng(5).uniform(-1, 1, (6, 7, 3)).astype(F)
    tv, g = tfops.total_variation(x)
    eps = 1e-3
    for idx in [(0, 0, 0), (3, 4, 1), (5, 6, 2)]:
        xp = x.copy(); xp[idx] += eps
        xm = x.copy(); xm[idx] -= eps
        num = (tfops.total_variation(xp)[0] - tfops.total_variation(xm)[0]) / (2 * eps)
        assert abs(num - g[idx]) < 2e-2


# ---- known answers derived by hand from the published TF / TFA algorithms (no TensorFlow needed) -------------------------
def test_scale_and_translate_known_answers_4_to_2_and_3_to_5():
    """tensorflow/core/kernels/image/scale_and_translate_op.cc, ComputeSpansCore with the triangle kernel, antialias=True.
    4 -> 2: kernel scale 2, sample positions 1.0 and 3.0, taps |(k + .5 - s) / 2| -> weights (3,3,1)/7 at 0..2 and
    (1,3,3)/7 at 1..3.   3 -> 5: kernel scale 1 (up-sampling), sample positions .3 .9 1.5 2.1 2.7 -> plain bilinear taps."""
    starts, w, n = tfops.compute_spans(2, 4)
    assert n == 4 and list(starts) == [0, 1]
    np.testing.assert_allclose(w[0, :3], np.array([3, 3, 1], F) / F(7), rtol=0, atol=1e-7)
    np.testing.assert_allclose(w[1, :3], np.array([1, 3, 3], F) / F(7), rtol=0, atol=1e-7)
    assert w[0, 3] == 0 and w[1, 3] == 0
    ramp = np.arange(4, dtype=F)
    x = np.broadcast_to(ramp[:, None, None], (4, 4, 3)).copy()
    np.testing.assert_allclose(tfops.aa_resize(x, 2, 2)[:, 0, 0], [5 / 7, 16 / 7], rtol=5e-7)       # float32 accumulation
    starts, w, n = tfops.compute_spans(5, 3)
    assert n == 3
    x3 = np.array([1.0, 10.0, 100.0], F)
    col = np.broadcast_to(x3[:, None, None], (3, 3, 3)).copy()
    got = tfops.aa_resize(col, 5, 5)[:, 0, 0]
    want = [1.0, 0.6 * 1 + 0.4 * 10, 10.0, 0.4 * 10 + 0.6 * 100, 100.0]
    np.testing.assert_allclose(got, want, rtol=3e-7)


def test_projective_transform_known_answer_90_degrees_on_a_ramp():
    """tfa.image.rotate(+90 deg) = ImageProjectiveTransformV3 with angles_to_projective_transforms: for a 3x3 image
    x_offset = ((W-1) - (cos (W-1) - sin (H-1))) / 2 = 2, y_offset = 0, so output(y, x) = input(row x, column 2 - y)."""
    img = np.arange(9, dtype=F).reshape(3, 3, 1)
    T = tfops.rotation_transform(F(0.0), F(1.0), 3)
    np.testing.assert_allclose(T, [0, -1, 2, 1, 0, 0, 0, 0], atol=0)
    out = tfops.projective_bilinear(img, T, -2.0)[..., 0]
    np.testing.assert_array_equal(out, np.array([[2, 5, 8], [1, 4, 7], [0, 3, 6]], F))
    # a half-pixel shift: bilinear blend with the fill value outside (CONSTANT fill mode)
    Tshift = np.array([1, 0, 0.5, 0, 1, 0, 0, 0], F)
    out = tfops.projective_bilinear(img, Tshift, -2.0)[..., 0]
    np.testing.assert_array_equal(out[0], np.array([0.5, 1.5, 0.5 * 2 + 0.5 * -2.0], F))


def test_projective_gradient_is_the_inverse_warp_with_zero_fill():
    """tensorflow/python/ops/image_ops.py, _image_projective_transform_v3_grad: the registered gradient of the op is the
    op itself applied to the incoming gradient with the INVERTED transform and fill 0 (not the scatter adjoint)."""
    D = 23
    rng = np.random.default_rng(6)
    T = tfops.rotation_transform(F(np.cos(0.3)), F(np.sin(0.3)), D, F(1e-4), F(-2e-4))
    for g in (np.eye(D, dtype=F)[:, :, None].repeat(3, 2), rng.normal(size=(D, D, 3)).astype(F)):
        got = tfops.projective_bilinear_grad(g, T)
        want = tfops.projective_bilinear(g, tfops.invert_transform(T), 0.0)
        np.testing.assert_array_equal(got, want)
    delta = np.zeros((D, D, 3), F)
    delta[11, 7] = 1.0
    back = tfops.projective_bilinear_grad(delta, T)
    assert back.sum() > 0 and (back > 0).sum() <= 4 * 3          # one delta spreads over at most four taps


def test_patcher_chain_gradient_against_finite_differences_without_rotation():
    """Everything in the backward chain except the rotate gradient (whose definition is pinned above): with angle 0 the
    warp is the identity sampling, so d(sum G * out)/d(patch) from the oracle's restated chain must agree with central
    differences of the oracle's forward (resize adjoint, noise / delta pass-through, inner and outer clips away from
    their kinks, print adjust, brightness-match mean term)."""
    from mladversarialobjectdetection_b200 import synth
    from oracle import patcher
    H, P = 96, 20
    bt = synth.make_batch(2, H, H, seed=41, max_boxes=2, min_boxes=2)
    bt.params["cos"], bt.params["sin"] = F(1.0), F(0.0)
    bt.params["delta"] = F(0.05)
    bt.images *= F(0.5)
    patch = (synth.make_patch(P, seed=41) * F(0.5)).astype(F)         # away from the clip limits
    bx, pr = bt.ragged()
    G = np.random.default_rng(42).normal(size=bt.images.shape).astype(F)

    def loss(p):
        out, _, _ = patcher.patcher_forward(p.astype(F), bt.images, bx, pr, bt.print_wb, 0.5)
        return float((out.astype(np.float64) * G).sum())

    _, _, states = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, 0.5)
    g = patcher.patcher_backward(G, patch, bt.print_wb, states, dtype=np.float64)
    rng = np.random.default_rng(43)
    eps = 2e-2
    for _ in range(4):
        d = rng.normal(size=patch.shape).astype(F)
        num = (loss(patch + F(eps) * d) - loss(patch - F(eps) * d)) / (2 * eps)
        ana = float((g * d).sum())
        assert abs(num - ana) <= 2e-2 * max(1.0, abs(ana)), (num, ana)


@pytest.mark.parametrize("deg,pa,pb", [(15.0, 0.0, 0.0), (-20.0, 0.0, 0.0), (7.0, 2e-4, -1.5e-4)])
def test_projective_matches_scipy_grid_constant_interpolation(deg, pa, pb):
    """Independent implementation of the same sampling rule: scipy.ndimage.map_coordinates(order=1, mode='grid-constant',
    cval=fill) is bilinear interpolation over an image extended by the constant on every side, i.e. taps outside the
    image READ AS THE FILL and are blended (ImageProjectiveTransformV3 BILINEAR + CONSTANT) -- the
    edge behaviour attacker.py:440's `< -1` test depends on.  Same float32 coordinates, float64 blend: agreement at
    float32 rounding level, and the set of elements below -1 is the same away from rounding ties."""
    from scipy import ndimage
    D = 97
    img = np.random.default_rng(5).uniform(-1, 1, (D, D, 3)).astype(F)
    th = deg * np.pi / 180
    T = tfops.rotation_transform(F(np.cos(th)), F(np.sin(th)), D, F(pa), F(pb))
    R = tfops.projective_bilinear(img, T, -2.0)
    ix, iy, _ = tfops.projective_coords(T, D, D)
    coords = np.stack([np.asarray(iy, dtype=np.float64), np.asarray(ix, dtype=np.float64)])
    S = np.stack([ndimage.map_coordinates(img[..., c].astype(np.float64), coords, order=1, mode="grid-constant", cval=-2.0)
                  for c in range(3)], -1)
    assert np.abs(R - S).max() < 2e-6
    clear = np.abs(S + 1.0) > 1e-5                      # away from the threshold the two roundings agree on the mask
    assert np.array_equal((R < -1)[clear], (S < -1)[clear])
    assert 0.03 < (R < -1).mean() < 0.25


@pytest.mark.parametrize("P,ps", [(100, 37), (100, 64), (100, 160), (30, 97), (64, 64)])
def test_aa_resize_matches_pillow_float_bilinear(P, ps):
    """Second independent implementation of the antialiased triangle resize: Pillow's Image.resize(BILINEAR) on a float
    ('F' mode) image widens the triangle by the down-sampling factor and normalises the taps exactly like
    ScaleAndTranslate(kernel 'triangle', antialias=True) (tf.image.resize's documentation names Pillow as its model).
    Pillow runs the horizontal pass first and keeps float64 taps and a float64 scale, TF a float32 scale and its float32
    reciprocal: agreement at the level of float32 rounding of the sample positions (measured: 0 ... 4.3e-6)."""
    PIL = pytest.importorskip("PIL.Image")
    x = np.random.default_rng(P * 1000 + ps).uniform(-1, 1, (P, P, 3)).astype(F)
    y = tfops.aa_resize(x, ps, ps)
    z = np.stack([np.asarray(PIL.fromarray(x[..., c], mode="F").resize((ps, ps), resample=PIL.BILINEAR)) for c in range(3)], -1)
    assert z.shape == y.shape
    assert np.abs(y - z).max() < 1e-5
