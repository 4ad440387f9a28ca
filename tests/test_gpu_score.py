"""GPU parity: fused sigmoid / arg-max-class / valid-box / per-image max and its dense gradient vs
the oracle (attacker.py:118-141,190-193; tf2/postprocess.py:104-156)."""
import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import ops
from oracle import objective as ob

pytestmark = pytest.mark.gpu
F = np.float32


def _heads(B, size, seed, person_boost=2.0, mu=-3.0):
    rng = np.random.default_rng(seed)
    fs = ob.feat_sizes(size, 7)[3:]
    cls = [rng.normal(mu, 2, (B, h, w, 810)).astype(F) for h, w in fs]
    box = [rng.normal(0, 0.5, (B, h, w, 36)).astype(F) for h, w in fs]
    for c in cls:
        c.reshape(B, -1, 90)[..., 0] += F(person_boost)
    return cls, box


def _run(cls, box, size, scale):
    anchors = ob.anchor_boxes(size)
    cls_t = [torch.from_numpy(c).cuda() for c in cls]
    box_t = [torch.from_numpy(b).cuda() for b in box]
    M, am, nc, ctx = ops.score_max_forward(cls_t, box_t, torch.from_numpy(anchors).cuda(), (size, size))
    sc = torch.tensor(scale, dtype=torch.float32, device="cuda")
    dcls, dscale, loss = ops.score_max_backward(ctx, sc)
    torch.cuda.synchronize()
    c, b = ob.merge_levels(cls, box, 90)
    post = ob.objective_forward(c, b, anchors, size, size, scale)
    return (M.cpu().numpy(), am.cpu().numpy(), nc.cpu().numpy(), [d.cpu().numpy() for d in dcls],
            float(dscale), float(loss), post, c)


def _guard_band(post, size):
    """Anchors whose validity test sits within float rounding of a threshold (expf vs np.exp)."""
    near = (np.abs(post["area"] - 100) < 1e-3) | (np.abs(post["bh"] - size) < 1e-3) | (np.abs(post["bw"] - size) < 1e-3)
    return int((near & (post["cls"] == 0)).sum())


@pytest.mark.parametrize("B,size,seed", [(3, 64, 1), (2, 512, 2), (2, 640, 3), (1, 1024, 4), (5, 96, 5)])
def test_score_max_forward_backward(B, size, seed):
    scale = 0.4
    cls, box = _heads(B, size, seed)
    M, am, nc, dcls, dscale, loss, post, c = _run(cls, box, size, scale)
    assert _guard_band(post, size) == 0, "test data sits on a validity threshold; change the seed"
    np.testing.assert_array_equal(nc, post["cand"].sum(1))                  # bit-exact candidate mask size
    np.testing.assert_allclose(M, post["max_scores"], rtol=0, atol=1e-6)
    for b in range(B):
        if post["has"][b]:
            sc = np.where(post["cand"][b], post["score"][b], -1)
            assert am[b] == int(np.argmax(sc))
        else:
            assert am[b] == -1 and M[b] == 0
    ref_d, ref_ds = ob.objective_backward(c, post, scale)
    ref_levels = ob.split_levels(ref_d, [x.shape for x in cls])
    for got, ref in zip(dcls, ref_levels):
        assert (got != 0).sum() == (ref != 0).sum()
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-9)
    assert abs(dscale - float(ref_ds)) < 1e-5
    assert abs(loss - float(post["loss"])) <= 1e-5 * max(1.0, abs(float(post["loss"])))


def test_score_max_no_candidates_and_ties():
    size, scale = 64, 0.4
    cls, box = _heads(3, size, 7)
    for c in cls:
        v = c.reshape(3, -1, 90)
        v[1, :, 0] = -50.0        # image 1: person is never arg-max
        v[2] = 0.25               # image 2: all logits equal -> all anchors tie, all classes tie
    M, am, nc, dcls, dscale, loss, post, c = _run(cls, box, size, scale)
    assert am[1] == -1 and M[1] == 0.0 and nc[1] == 0
    ref_d, ref_ds = ob.objective_backward(c, post, scale)
    ref_levels = ob.split_levels(ref_d, [x.shape for x in cls])
    nz = 0
    for got, ref in zip(dcls, ref_levels):
        assert not got[1].any()
        assert (got[2] != 0).sum() == (ref[2] != 0).sum()
        nz += int((got[2] != 0).sum())
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-12)
    assert nz == int(post["cand"][2].sum()) * 90 > 0       # every valid anchor ties, all 90 classes tie
    assert abs(dscale - float(ref_ds)) < 1e-5


def test_score_max_rejects_bad_shapes():
    cls, box = _heads(1, 64, 9)
    anchors = torch.from_numpy(ob.anchor_boxes(64)).cuda()
    cls_t = [torch.from_numpy(c).cuda() for c in cls]
    box_t = [torch.from_numpy(b).cuda() for b in box]
    with pytest.raises(ValueError):
        ops.score_max_forward(cls_t, box_t, anchors[:-1], (64, 64))
    with pytest.raises(RuntimeError, match="CUDA only"):
        ops.score_max_forward([c.cpu() for c in cls_t], box_t, anchors, (64, 64))
