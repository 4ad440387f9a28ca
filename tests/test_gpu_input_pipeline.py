"""Input pipeline kernels (csrc/input_pipeline.cu) through the C ABI against the oracle and the reference fixtures
(train_data_generator.py:55-75, 201-226).  Bar: float32 outputs equal up to 1 ulp on isolated elements
(the float64 sums / products may associate differently), documented in oracle/input_pipeline.py."""
import os

import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import ops, train_data_generator as tdg
from oracle import input_pipeline as ip

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NORMS = {"a": (127.0, 128.0), "b": ([123.675, 116.28, 103.53], [58.395, 57.12, 57.375])}


def test_letterbox_matches_reference_fixtures_exactly_sized_batch():
    g = np.load(os.path.join(GOLD, "map_fn.npz"))
    n = int(g["n"])
    frames = [g[f"frame{k}"] for k in range(n)]
    for tag, (mean, std) in NORMS.items():
        seq = tdg.DataSequence("", (64, 64), mean, std, file_list=["unused"])
        out, sums = seq.map_batch(frames)
        out = out.cpu().numpy()
        for k in range(n):
            want = g[f"out{k}{tag}"]
            np.testing.assert_allclose(out[k], want, rtol=0, atol=1.2e-7)
            assert (out[k] != want).mean() < 1e-4
            np.testing.assert_array_equal(out[k], ip.map_fn(frames[k], (64, 64), mean, std))   # same order as the oracle
        np.testing.assert_allclose(sums.cpu().numpy(), out.astype(np.float64).sum(axis=(1, 2)), rtol=1e-12, atol=1e-9)


def test_letterbox_large_ragged_batch_more_than_one_launch():
    rng = np.random.default_rng(9)
    frames = [rng.integers(0, 256, size=(int(rng.integers(20, 300)), int(rng.integers(20, 300)), 3), dtype=np.uint8)
              for _ in range(70)]                                   # > 64 frames: two launches
    frames[3] = rng.integers(0, 256, size=(96, 96, 3), dtype=np.uint8)        # same size: copy path
    frames[5] = rng.integers(0, 256, size=(192, 192, 3), dtype=np.uint8)      # exact 2x
    seq = tdg.DataSequence("", (96, 96), 127.0, 128.0, file_list=["unused"])
    out, _ = seq.map_batch(frames)
    out = out.cpu().numpy()
    for k, f in enumerate(frames):
        np.testing.assert_array_equal(out[k], ip.map_fn(f, (96, 96), 127.0, 128.0))


def test_augment_matches_oracle():
    rng = np.random.default_rng(10)
    x = rng.uniform(-1, 1, (5, 40, 56, 3)).astype(np.float32)
    flip = np.array([1, 0, 1, 1, 0], np.uint8)
    for contrast, delta in [(0.83, 0.17), (1.2, -0.2), (1.0, 0.0)]:
        want = ip.augment(x, flip, contrast, delta)
        got = ops.augment_batch(torch.from_numpy(x).cuda(), torch.from_numpy(flip).cuda(), contrast, delta).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=0, atol=2.4e-7)
        assert (got != want).mean() < 1e-2
    got = ops.augment_batch(torch.from_numpy(x).cuda(), None, 1.0, 0.0).cpu().numpy()
    np.testing.assert_allclose(got, x, atol=2.4e-7)


def test_full_size_frames_properties():
    """config-2 size: 512x512 output from 480x640 frames -- padding rows are exactly zero, values stay in the
    standardised range, a second pass over the same frames is idempotent (bit-identical)."""
    rng = np.random.default_rng(11)
    frames = [rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8) for _ in range(8)]
    seq = tdg.DataSequence("", (512, 512), 127.0, 128.0, file_list=["unused"])
    out, sums = seq.map_batch(frames)
    out2, _ = seq.map_batch(frames)
    assert torch.equal(out, out2)
    assert not out[:, 384:].any() and float(out.max()) <= 1.0 and float(out.min()) >= -127.0 / 128.0
    aug = tdg.augment(out, np.random.default_rng(1), sums=sums)
    assert aug.shape == out.shape and float(aug.max()) <= 1.0 and float(aug.min()) >= -1.0


def test_errors():
    with pytest.raises(RuntimeError, match="CUDA only"):
        ops.letterbox_normalize([torch.zeros((4, 4, 3), dtype=torch.uint8)], (8, 8), 127.0, 128.0)
    with pytest.raises(RuntimeError, match="scales to"):
        ops.letterbox_normalize([torch.zeros((1, 4000, 3), dtype=torch.uint8, device="cuda")], (8, 8), 127.0, 128.0)


def test_augment_scalar_path_odd_width_and_channel_sums():
    """W % 4 != 0 takes the scalar kernel; eot_channel_sums feeds the contrast means when the batch did not come from
    the letter-box kernel."""
    rng = np.random.default_rng(12)
    x = rng.uniform(-1, 1, (3, 17, 57, 3)).astype(np.float32)
    flip = np.array([0, 1, 1], np.uint8)
    xt = torch.from_numpy(x).cuda()
    sums = ops.channel_sums(xt)
    np.testing.assert_allclose(sums.cpu().numpy(), x.astype(np.float64).sum(axis=(1, 2)), rtol=1e-12, atol=1e-9)
    got = ops.augment_batch(xt, torch.from_numpy(flip).cuda(), 0.9, -0.1, sums=sums).cpu().numpy()
    np.testing.assert_allclose(got, ip.augment(x, flip, 0.9, -0.1), rtol=0, atol=2.4e-7)
