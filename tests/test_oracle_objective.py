"""Oracle self-checks for anchors / pre_nms / person-score objective."""
import numpy as np
import pytest
import torch

from oracle import objective as ob

F = np.float32


def test_first_anchor_known_answer():
    # tf2/postprocess_test.py:27-35,229: min_level 1, max_level 2, 1 scale, ratio 1, anchor_scale 1, image 8
    # (normalised, centre-size encoding of decode_anchors_to_centersize, tf2/anchors.py:61-80)
    a = ob.anchor_boxes(8, 1, 2, 1, (1.0,), 1.0)
    n = a[0] / 8.0
    cs = [(n[0] + n[2]) / 2, (n[1] + n[3]) / 2, n[2] - n[0], n[3] - n[1]]
    np.testing.assert_allclose(cs, [0.125, 0.125, 0.25, 0.25])
    assert a.shape == (16 + 4, 4)


@pytest.mark.parametrize("size,A,feats", [(512, 49104, [64, 32, 16, 8, 4]), (640, 76725, [80, 40, 20, 10, 5]),
                                          (1024, 196416, [128, 64, 32, 16, 8])])
def test_anchor_counts_and_feature_sizes(size, A, feats):
    assert [f[0] for f in ob.feat_sizes(size, 7)[3:]] == feats
    a = ob.anchor_boxes(size)
    assert a.shape == (A, 4) and a.dtype == F
    # level-3 first location: stride 8, centre 4, base 32 -> aspect 1 box is 32x32
    np.testing.assert_allclose(a[0], [4 - 16, 4 - 16, 4 + 16, 4 + 16])
    # aspect 2.0: wider than tall
    assert (a[1, 3] - a[1, 1]) > (a[1, 2] - a[1, 0])


def _random_heads(B, size, rng, person_boost=0.0):
    fs = ob.feat_sizes(size, 7)[3:]
    cls = [rng.normal(-3, 2, (B, h, w, 9 * 90)).astype(F) for h, w in fs]
    box = [rng.normal(0, 0.5, (B, h, w, 9 * 4)).astype(F) for h, w in fs]
    for c in cls:
        c.reshape(B, -1, 90)[..., 0] += person_boost
    return cls, box


def test_decode_zero_regression_returns_anchor():
    a = ob.anchor_boxes(64)
    d = ob.decode(np.zeros((1, len(a), 4), F), a[None])
    np.testing.assert_allclose(d[0], a, atol=1e-5)


def test_objective_forward_semantics():
    rng = np.random.default_rng(0)
    cls, box = _random_heads(2, 64, rng, person_boost=2.0)
    anchors = ob.anchor_boxes(64)
    c, b = ob.merge_levels(cls, box, 90)
    assert c.shape == (2, len(anchors), 90)
    c[1, :, 0] = -50.0                      # image 1: person never arg-max -> no candidate -> clamp at 0
    post = ob.objective_forward(c, b, anchors, 64, 64, 0.4)
    assert post["has"][0] and not post["has"][1]
    assert post["max_scores"][1] == 0.0
    cand = post["cand"][0]
    assert post["max_scores"][0] == post["score"][0][cand].max()
    assert (post["cls"][0][cand] == 0).all() and (post["area"][0][cand] > 100).all()
    assert (post["bh"][0][cand] <= 64).all() and (post["bw"][0][cand] <= 64).all()


def test_objective_backward_matches_autograd_including_ties():
    rng = np.random.default_rng(1)
    cls, box = _random_heads(3, 64, rng, person_boost=3.0)
    anchors = ob.anchor_boxes(64)
    c, b = ob.merge_levels(cls, box, 90)
    c = c.copy()
    # image 2: all logits equal -> every valid anchor ties, and all 90 classes tie
    c[2] = 0.25
    scale = 0.4
    post = ob.objective_forward(c, b, anchors, 64, 64, scale)
    dcls, dscale = ob.objective_backward(c, post, scale, dtype=np.float64)
    ct = torch.tensor(c.astype(np.float64), requires_grad=True)
    st = torch.tensor(scale, dtype=torch.float64, requires_grad=True)
    logit = ct.amax(dim=-1)                 # amax splits gradients equally among ties (like TF Max grad)
    score = torch.sigmoid(logit)
    cand = torch.from_numpy(post["cand"])
    masked = torch.where(cand, score, torch.full_like(score, -1.0))
    M = torch.clamp(masked.amax(dim=1), min=0.0)
    loss = (M ** 2 + (M - st) ** 2).sum()
    loss.backward()
    # SigmoidGrad is dy*y*(1-y) on the float32 y (TF); autograd here uses a float64 y -> (1-y) cancellation
    np.testing.assert_allclose(dcls, ct.grad.numpy(), rtol=5e-4, atol=1e-12)
    assert abs(dscale - st.grad.item()) < 1e-6
    assert abs(float(post["loss"]) - loss.item()) < 1e-5
    assert (dcls[2] != 0).sum() == post["cand"][2].sum() * 90


def test_split_levels_roundtrip():
    rng = np.random.default_rng(2)
    cls, box = _random_heads(2, 64, rng)
    c, _ = ob.merge_levels(cls, box, 90)
    parts = ob.split_levels(c, [x.shape for x in cls])
    for p, x in zip(parts, cls):
        np.testing.assert_array_equal(p, x)
