"""Oracle of the input pipeline (oracle/input_pipeline.py) against the REFERENCE's own `DataSequence._map_fn`
(fixtures tests/golden/map_fn.npz, generated in the build container by importing /root/reference/train_data_generator.py)."""
import os

import numpy as np
import pytest

from oracle import input_pipeline as ip

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NORMS = {"a": (127.0, 128.0), "b": ([123.675, 116.28, 103.53], [58.395, 57.12, 57.375])}


def test_map_fn_matches_reference_fixtures():
    g = np.load(os.path.join(GOLD, "map_fn.npz"))
    for k in range(int(g["n"])):
        frame = g[f"frame{k}"]
        for tag, (mean, std) in NORMS.items():
            want = g[f"out{k}{tag}"]
            got = ip.map_fn(frame, (64, 64), mean, std)
            assert got.dtype == np.float32 and got.shape == (64, 64, 3)
            # bar: identical float32 up to 1 ulp on isolated elements (IPP's float64 evaluation order is its own)
            np.testing.assert_allclose(got, want, rtol=0, atol=1.2e-7)
            assert (got != want).mean() < 1e-4


def test_resize_against_cv2_when_installed():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for _ in range(12):
        sh, sw, dh, dw = (int(v) for v in rng.integers(2, 300, 4))
        img = (rng.integers(0, 256, size=(sh, sw, 3)).astype(np.float64) - 127.0) / 128.0
        want = cv2.resize(img, [dw, dh]).astype(np.float32)
        got = ip.cv2_resize_linear_f64(img, dw, dh).astype(np.float32)
        np.testing.assert_allclose(got, want, rtol=0, atol=1.2e-7)


def test_scaled_size_and_padding():
    assert ip.scaled_size(480, 640, 512, 512) == (384, 512)
    assert ip.scaled_size(640, 480, 512, 512) == (512, 384)
    out = ip.map_fn(np.full((10, 20, 3), 255, np.uint8), (16, 16), 127.0, 128.0)
    assert np.all(out[:8, :16] == np.float32(1.0)) and not out[8:].any()


def test_augment_semantics():
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, (2, 6, 5, 3)).astype(np.float32)
    y = ip.augment(x, np.array([1, 0]), 1.0, 0.0)
    np.testing.assert_allclose(y[0], x[0, :, ::-1], atol=1e-6)
    np.testing.assert_allclose(y[1], x[1], atol=1e-6)
    y = ip.augment(x, np.array([0, 0]), 0.8, 0.5)
    assert y.max() <= 1.0 and y.min() >= -1.0
    m = x.mean(axis=(1, 2), keepdims=True)
    np.testing.assert_allclose(y, np.clip((x - m) * 0.8 + m + 0.5, -1, 1), atol=1e-6)
