"""Oracle of the uint8 inference twin (oracle/adv_patch_u8.py) against the REFERENCE's own
`adv_patch.AdversarialPatch.add_adv_to_img` (fixtures tests/golden/adv_patch_u8.npz) and, where OpenCV is
installed, against cv2 itself.  Everything here is bit-exact."""
import os

import numpy as np
import pytest

from oracle import adv_patch_u8 as o

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cases():
    g = np.load(os.path.join(GOLD, "adv_patch_u8.npz"))
    P = int(g["P"])
    for k in range(int(g["n"])):
        boxes = [tuple(b) for b in g[f"boxes{k}"]]
        yield dict(P=P, raw=g[f"raw{k}"], printed=g[f"printed{k}"], frame=g[f"frame{k}"], boxes=boxes,
                   scale=float(g[f"scale{k}"]), result=g[f"result{k}"], noises=[g[f"noise{k}_{i}"] for i in range(len(boxes))])


def test_print_patch_and_full_paste_match_reference_fixtures():
    n = 0
    for c in _cases():
        printed = o.print_patch(c["raw"])
        np.testing.assert_array_equal(printed, c["printed"])
        got = o.add_adv_to_img(c["frame"], c["boxes"], printed, (c["P"], c["P"]), c["scale"], c["noises"])
        np.testing.assert_array_equal(got, c["result"])
        assert (got != c["frame"]).any()
        n += 1
    assert n == 5          # the last case takes the INTER_CUBIC branch (boxes longer than twice the texture side)


def test_create_matches_reference_create_fixture():
    rows = np.load(os.path.join(GOLD, "adv_patch_create.npz"))["rows"]
    for H, W, scale, ymin, xmin, ymax, xmax, ry, rx, rph, rpw in rows:
        assert o.create(int(H), int(W), (ymin, xmin, ymax, xmax), scale) == (int(ry), int(rx), int(rph), int(rpw))


def test_colour_conversions_against_cv2_all_colours():
    cv2 = pytest.importorskip("cv2")
    allc = np.stack(np.meshgrid(np.arange(0, 256, 3), np.arange(256), np.arange(256), indexing="ij"), -1)
    allc = allc.reshape(-1, 256, 3).astype(np.uint8)
    np.testing.assert_array_equal(o.rgb2yuv(allc), cv2.cvtColor(allc, cv2.COLOR_RGB2YUV))
    np.testing.assert_array_equal(o.yuv2rgb(allc), cv2.cvtColor(allc, cv2.COLOR_YUV2RGB))


def test_resizes_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, size=(96, 96, 3), dtype=np.uint8)
    for d in (7, 24, 31, 32, 48, 50, 77, 95):                       # integer ratios (fast paths) and general ones
        np.testing.assert_array_equal(o.resize_area_u8(img, d, d), cv2.resize(img, (d, d), interpolation=cv2.INTER_AREA))
    for _ in range(25):
        sh, sw, dh, dw = (int(v) for v in rng.integers(6, 200, 4))
        im = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
        np.testing.assert_array_equal(o.resize_linear_u8(im, dw, dh), cv2.resize(im, (dw, dh)))
    im = rng.integers(0, 256, size=(64, 128, 3), dtype=np.uint8)
    np.testing.assert_array_equal(o.resize_linear_u8(im, 64, 32), cv2.resize(im, (64, 32)))   # exact 2x


def test_bicubic_upsampling_against_cv2():
    """INTER_CUBIC on 8-bit data: bit-identical to OpenCV's own kernel (IPP off); the IPP build of the same call (what a
    pip wheel runs by default) is a closed-source float evaluation that differs by at most one grey level."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(13)
    had = cv2.ipp.useIPP()
    try:
        for _ in range(12):
            sh, sw = (int(v) for v in rng.integers(8, 120, 2))
            dh, dw = sh + int(rng.integers(1, 200)), sw + int(rng.integers(1, 200))
            im = rng.integers(0, 256, size=(sh, sw, 3), dtype=np.uint8)
            got = o.resize_cubic_u8(im, dw, dh)
            cv2.ipp.setUseIPP(False)
            np.testing.assert_array_equal(got, cv2.resize(im, (dw, dh), interpolation=cv2.INTER_CUBIC))
            if had:
                cv2.ipp.setUseIPP(True)
                ipp = cv2.resize(im, (dw, dh), interpolation=cv2.INTER_CUBIC)
                assert np.abs(got.astype(int) - ipp.astype(int)).max() <= 1
    finally:
        cv2.ipp.setUseIPP(had)
