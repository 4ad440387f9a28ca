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
