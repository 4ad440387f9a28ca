"""BASELINE.json's full single-GPU size (config 2: 64 images of 512x512, 100x100 patch, 1..8 boxes per image, D0 heads):
the oracle would take minutes, so the CUDA path is checked through size-independent properties -- locality of the
paste, determinism, shard invariance (what data-parallel sharding relies on), linearity of the adjoint in dL/dimage,
and the defining properties of the score reduction and of its sparse gradient."""
import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import anchors as anchors_mod, ops, synth
from mladversarialobjectdetection_b200.anchors import feature_sizes
from tests._util import to_device

pytestmark = pytest.mark.gpu
B, H, P = 64, 512, 100


@pytest.fixture(scope="module")
def batch():
    bt = synth.make_batch(B, H, H, seed=1234, max_boxes=8)
    d = to_device(bt)
    d["patch"] = torch.from_numpy(synth.make_patch(P)).cuda()
    d["scale"] = torch.tensor(0.4, device="cuda")
    d["bt"] = bt
    return d


def _fwd(d, lo=0, hi=B):
    off = d["offsets"][lo:hi + 1]
    j0, j1 = int(off[0]), int(off[-1])
    return ops.apply_forward(d["patch"], d["scale"], d["images"][lo:hi], d["boxes"][j0:j1], (off - off[0]).contiguous(),
                             d["params"][j0:j1], d["print_wb"][lo:hi], ops.PatchGeometry())


def test_forward_locality_determinism_and_range(batch):
    out, _, ctx = _fwd(batch)
    ops.check_workspace(ctx)
    out2, _, _ = _fwd(batch)
    assert torch.equal(out, out2)                                           # no atomics / ordering in the composite
    geo = ops.box_geometry(tuple(batch["images"].shape), P, batch["boxes"], batch["offsets"], batch["params"],
                           batch["scale"]).cpu().numpy()
    inside = torch.zeros((B, H, H), dtype=torch.bool, device="cuda")
    off = batch["offsets"].cpu().numpy()
    for b in range(B):
        for j in range(off[b], off[b + 1]):
            y0, x0, ps, d, _, _, valid = geo[j][:7]
            if valid:
                inside[b, y0:y0 + d, x0:x0 + d] = True
    changed = (out != batch["images"]).any(-1)
    assert not (changed & ~inside).any()                                    # nothing outside the windows moves
    assert changed.float().mean() > 0.1 and float(out.max()) <= 1.0 and float(out.min()) >= -1.0


def test_forward_and_backward_are_shard_invariant(batch):
    """images [0,32) and [32,64) patched separately == the full batch; dL/dpatch adds up (the NCCL sum all-reduce)."""
    out, _, ctx = _fwd(batch)
    G = torch.randn_like(out)
    g_full = ops.apply_backward(ctx, G)
    g_sum = torch.zeros_like(g_full)
    for lo, hi in ((0, 32), (32, 64)):
        o, _, c = _fwd(batch, lo, hi)
        assert torch.equal(o, out[lo:hi])
        g_sum += ops.apply_backward(c, G[lo:hi].contiguous())
    rel = float((g_sum - g_full).norm() / g_full.norm())
    assert rel < 1e-5


def test_backward_is_linear_in_the_upstream_gradient_and_local(batch):
    out, _, ctx = _fwd(batch)
    G1, G2 = torch.randn_like(out), torch.randn_like(out)
    g1, g2 = ops.apply_backward(ctx, G1).clone(), ops.apply_backward(ctx, G2).clone()
    g12 = ops.apply_backward(ctx, (0.5 * G1 - 2.0 * G2).contiguous())
    assert float((g12 - (0.5 * g1 - 2.0 * g2)).norm() / g12.norm()) < 1e-5
    changed = (out != batch["images"]).any(-1, keepdim=True)
    g1_local = ops.apply_backward(ctx, (G1 * changed).contiguous())          # only pasted pixels carry gradient to the patch
    assert float((g1_local - g1).norm() / g1.norm()) < 1e-5
    assert torch.equal(ops.apply_backward(ctx, G1), g1)                      # deterministic reduction order


def test_score_reduction_properties():
    rng = np.random.default_rng(3)
    fs = feature_sizes((H, H), 7)[3:]
    dev = "cuda"
    cls = [torch.randn(B, h, w, 810, device=dev) * 1.5 - 4.6 for h, w in fs]
    box = [torch.randn(B, h, w, 36, device=dev) * 0.3 for h, w in fs]
    anc = torch.from_numpy(anchors_mod.anchor_table((H, H))).to(dev)
    M, argmax, ncand, ctx = ops.score_max_forward(cls, box, anc, (H, H))
    cand = ops.score_candidate_view(ctx)
    assert torch.equal(M, torch.clamp(cand.max(dim=1).values, min=0.0))      # maximum(reduce_max(ragged), 0)
    assert torch.equal((cand >= 0).sum(1).int(), ncand)
    rows = torch.arange(B, device=dev)
    has = ncand > 0
    assert torch.equal(cand[rows[has], argmax[has].long()], M[has])
    assert bool((cand[:, :] <= M[:, None]).all())
    # the candidate of the arg-max anchor is a person arg-max: logit 0 is the row maximum
    allc = torch.cat([c.reshape(B, -1, 90) for c in cls], 1)
    top = allc[rows[has], argmax[has].long()]
    assert torch.equal(top[:, 0], top.max(dim=1).values)
    np.testing.assert_allclose(M[has].cpu().numpy(), torch.sigmoid(top[:, 0].double()).float().cpu().numpy(), rtol=0, atol=1e-7)
    # sparse gradient: one anchor per image (no ties with continuous logits), dL/dlogit = (4M - 2 scale) * M (1 - M)
    scale = torch.tensor(0.4, device=dev)
    dcls, dscale, loss = ops.score_max_backward(ctx, scale)
    dall = torch.cat([c.reshape(B, -1, 90) for c in dcls], 1)
    assert torch.equal((dall != 0).flatten(1).sum(1) > 0, has) and int((dall != 0).sum()) == int(has.sum())
    want = ((4 * M - 2 * scale) * M * (1 - M))[has]
    np.testing.assert_allclose(dall[rows[has], argmax[has].long(), 0].cpu().numpy(), want.cpu().numpy(), rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(float(dscale), float((-2 * (M - scale)).sum()), rtol=1e-5)
    np.testing.assert_allclose(float(loss), float((M * M + (M - scale) ** 2).sum()), rtol=1e-5)
    # shard invariance of the objective
    M2, _, _, _ = ops.score_max_forward([c[32:] for c in cls], [b[32:] for b in box], anc, (H, H))
    assert torch.equal(M2, M[32:])
