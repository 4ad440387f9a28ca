"""GPU: the whole attack step through the reference-shaped Python API (PatchAttacker.train_step) vs the
oracle port of the same step, with the SAME victim weights (GPU vs CPU convolutions differ in the last bits,
so the step-level bars are looser than the kernel-level ones in test_gpu_forward/backward/score)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from mladversarialobjectdetection_b200 import ops, synth, victim
from mladversarialobjectdetection_b200.attack_detection import Masker
from mladversarialobjectdetection_b200.attacker import PatchAttacker, Patcher
from mladversarialobjectdetection_b200.brightness_matcher import BrightnessMatcher
from mladversarialobjectdetection_b200.ragged import RaggedBoxes
from oracle import objective, patcher, step as ostep

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_train_step_matches_oracle_step(no_tf32):
    H, P, B = 128, 32, 3
    model = victim.get_victim_model("efficientdet-d0", device="cuda", image_size=H, seed=5)
    # make the person class competitive so that every image has candidates
    model.class_net.out_pw.bias.data.view(9, 90)[:, 0] += 6.0
    cpu_model = victim.get_victim_model("efficientdet-d0", device="cpu", image_size=H, seed=5)
    cpu_model.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    bt = synth.make_batch(B, H, H, seed=77, max_boxes=3, min_boxes=1)
    att = PatchAttacker(model, patch_size=P, device="cuda", seed=3)
    att.compile(learning_rate=1e-2)
    patch0 = att._patch.cpu().numpy().copy()
    images = torch.from_numpy(bt.images).cuda()
    boxes = RaggedBoxes(torch.from_numpy(bt.boxes).cuda(), torch.from_numpy(bt.offsets).cuda())
    tr = (ops.params_to_tensor(bt.params, "cuda"), torch.from_numpy(bt.print_wb).cuda())
    dscale, gpatch = att.call(images, training=True, boxes=boxes, transforms=tr)
    torch.cuda.synchronize()
    bx, pr = bt.ragged()
    ref = ostep.attack_step(cpu_model, patch0, 0.4, bt.images, bx, pr, bt.print_wb, objective.anchor_boxes(H),
                            first_pass=False)
    M = att._last["max_scores"].cpu().numpy()
    assert (ref["max_scores"] > 0).all()
    np.testing.assert_allclose(M, ref["max_scores"], atol=2e-4)
    assert abs(float(dscale) - float(ref["dscale"])) < 2e-3
    g = gpatch.cpu().numpy()
    gref = ref["grad_patch"] - F(1e-5) * __import__("oracle.tfops", fromlist=["x"]).total_variation(patch0)[1]
    rel = np.linalg.norm(g - gref) / np.linalg.norm(gref)
    assert rel < 2e-2, rel
    # full train_step: metrics + Adam + constraints
    att2 = PatchAttacker(model, patch_size=P, device="cuda", seed=3)
    m = att2.train_step(images, boxes=boxes, transforms=tr)
    torch.cuda.synchronize()
    assert abs(float(m["loss"]) - float(ref["loss"])) < 5e-3 * max(1.0, float(ref["loss"]))
    assert abs(float(m["tv_loss"]) - float(ref["tv"])) / float(ref["tv"]) < 1e-5
    newp = att2._patch.cpu().numpy()
    assert np.abs(newp).max() <= 1.0
    moved = np.abs(newp - patch0)
    assert 0.009 < np.median(moved[moved > 0]) <= 0.0101          # Adam's first step is lr * sign-like
    assert 0.0 <= float(att2._scale_regressor) <= 1.0 and float(att2._scale_regressor) != 0.4


def test_patcher_layer_signature_and_internal_sampler():
    H, P = 160, 40
    bt = synth.make_batch(2, H, H, seed=78, max_boxes=3)
    patch = torch.from_numpy(synth.make_patch(P)).cuda()
    scale = torch.tensor(0.4, device="cuda")
    layer = Patcher(patch, scale, min_patch_area=4, name="Patcher", seed=11)
    images = torch.from_numpy(bt.images).cuda()
    rows = [bt.boxes_of(b) for b in range(2)]
    out1 = layer([rows, images])                     # ragged python rows accepted, transforms drawn on device
    out2 = layer([rows, images])
    torch.cuda.synchronize()
    assert out1.shape == images.shape and not torch.equal(out1, out2)        # fresh transforms per call
    assert not torch.equal(out1, images)
    layer2 = Patcher(patch, scale, seed=11)
    assert torch.equal(layer2([rows, images]), out1)                         # same seed -> same draw
    # replay through the oracle with the sampler's draw
    s = layer2.sampler
    boxes = RaggedBoxes.from_rows(rows, "cuda")
    prm = s.box_params(0, 0, boxes.row_splits, boxes.values.shape[0]).cpu().numpy().view(synth.BOX_PARAMS).reshape(-1)
    wb = s.print_wb(0, 0, 2, "cuda").cpu().numpy()
    off = bt.offsets
    ref, _, _ = patcher.patcher_forward(patch.cpu().numpy(), bt.images, rows, [prm[off[b]:off[b + 1]] for b in range(2)], wb, 0.4)
    np.testing.assert_array_equal(out1.cpu().numpy(), ref)


def test_sharded_draw_equals_single_rank_draw():
    H, P, B = 128, 24, 4
    bt = synth.make_batch(B, H, H, seed=79, max_boxes=3)
    patch = torch.from_numpy(synth.make_patch(P)).cuda()
    scale = torch.tensor(0.4, device="cuda")
    images = torch.from_numpy(bt.images).cuda()
    full = RaggedBoxes(torch.from_numpy(bt.boxes).cuda(), torch.from_numpy(bt.offsets).cuda())
    whole = Patcher(patch, scale, seed=5)([full, images])
    parts = []
    for r in range(2):
        lay = Patcher(patch, scale, seed=5)
        lay.first_image = r * 2
        parts.append(lay([full.slice_rows(r * 2, r * 2 + 2), images[r * 2:r * 2 + 2]]))
    assert torch.equal(torch.cat(parts), whole)


def test_brightness_matcher_layer():
    rng = np.random.default_rng(80)
    src = rng.uniform(-1.3, 1.3, (37, 53, 3)).astype(F)      # not square, slightly out of range: no clip before rescale
    tgt = rng.uniform(-1, 1, (90, 70, 3)).astype(F)
    out = BrightnessMatcher(name="Brightness_Matcher")((torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()))
    ms = patcher.MatchState  # noqa
    from oracle import tfops
    s = (src + F(1)) * tfops.C_127_255
    yuv = tfops.dot3(s, tfops.RGB2YUV)
    mu_s, mu_t = tfops.mean_f64(yuv[..., 0]), patcher.image_mean_y(tgt)
    yp = np.clip((yuv[..., 0] - mu_s) + mu_t, 0, 1)
    rgb = tfops.dot3(np.stack([yp, yuv[..., 1], yuv[..., 2]], -1), tfops.YUV2RGB)
    ref = np.clip(rgb, 0, 1) * tfops.C_255_127 - F(1)
    np.testing.assert_array_equal(out.cpu().numpy(), ref.astype(F))


def test_masker_layer_outputs_and_eval_branch():
    H = 320
    bt = synth.make_batch(3, H, H, seed=81, max_boxes=3)
    images = torch.from_numpy(bt.images).cuda()
    boxes = RaggedBoxes(torch.from_numpy(bt.boxes).cuda(), torch.from_numpy(bt.offsets).cuda())
    patch = torch.from_numpy(synth.make_patch(64)).cuda()
    scale = torch.tensor(0.4, device="cuda")
    mk = Masker(patch, scale, seed=2)
    out, mask = mk([boxes, images], training=True)
    torch.cuda.synchronize()
    assert out.shape == images.shape == mask.shape
    changed = (out != images).any(-1)
    assert changed.any() and torch.equal(mask[changed], (images - out)[changed])
    assert float(mask[~changed].abs().max()) == 0.0
    out_e, mask_e = mk([boxes, images], training=False)        # learned patch, tolerance 0, shared scale
    assert not torch.equal(out_e, out)


def test_save_and_resume_weights_roundtrip():
    model = victim.get_victim_model("efficientdet-d0", device="cuda", image_size=128, seed=1)
    att = PatchAttacker(model, patch_size=24, device="cuda", seed=9)
    d = os.path.join(tempfile.mkdtemp(), "ckpt")
    att.save_weights(d)
    assert sorted(os.listdir(d)) == ["patch.png", "patch.tiff", "scale.txt"]
    att2 = PatchAttacker(model, initial_patch=d, device="cuda")
    assert torch.equal(att2._patch, att._patch) and abs(float(att2._scale_regressor) - 0.4) < 1e-7
    with pytest.raises(FileExistsError):
        att.save_weights(d)                                     # os.makedirs without exist_ok (attacker.py:334)


def test_cuda_graph_replay_equals_eager_step(no_tf32):
    H, P, B = 128, 32, 3
    model = victim.get_victim_model("efficientdet-d0", device="cuda", image_size=H, seed=5)
    model.class_net.out_pw.bias.data.view(9, 90)[:, 0] += 6.0
    bt = synth.make_batch(B, H, H, seed=79, max_boxes=3, min_boxes=1)
    images = torch.from_numpy(bt.images).cuda()
    boxes = RaggedBoxes(torch.from_numpy(bt.boxes).cuda(), torch.from_numpy(bt.offsets).cuda())
    tr = (ops.params_to_tensor(bt.params, "cuda"), torch.from_numpy(bt.print_wb).cuda())
    res = []
    for graphs in (False, True):
        att = PatchAttacker(model, patch_size=P, device="cuda", seed=3, cuda_graphs=graphs)
        att.compile(learning_rate=1e-2)
        for _ in range(3):                                  # several steps: the replayed graph must see the updated patch / scale
            m = att.train_step(images, boxes=boxes, transforms=tr)
        torch.cuda.synchronize()
        res.append((att._patch.clone(), float(att._scale_regressor), float(m["loss"])))
    # cuDNN picks algorithms independently for the two runs: equal up to float32 reduction order
    assert torch.allclose(res[0][0], res[1][0], atol=2e-3)
    assert (res[0][0] - res[1][0]).abs().mean() < 2e-4
    assert abs(res[0][1] - res[1][1]) < 1e-4
    assert abs(res[0][2] - res[1][2]) < 1e-3 * max(1.0, abs(res[0][2]))


def test_fused_bias_silu_epilogue_matches_torch_ops():
    """victim.conv_bias_act (cuDNN conv + libeotpatch's one-pass bias/SiLU epilogue) vs conv2d(bias) + silu of torch,
    forward and input gradient, channels_last, with and without activation; and the whole victim with the switch off."""
    torch.manual_seed(0)
    x = torch.randn(4, 16, 24, 20, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(32, 16, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last) * 0.1
    b = torch.randn(32, device="cuda")
    for act in (True, False):
        xa = x.clone().requires_grad_(True)
        ya = victim.conv_bias_act(xa, w, b, 1, 1, 1, 1, act)
        xb = x.clone().requires_grad_(True)
        yb = torch.nn.functional.conv2d(xb, w, b, 1, 1)
        yb = torch.nn.functional.silu(yb) if act else yb
        g = torch.randn_like(yb)
        ya.backward(g)
        yb.backward(g)
        torch.testing.assert_close(ya, yb, rtol=2e-6, atol=2e-6)
        torch.testing.assert_close(xa.grad, xb.grad, rtol=1e-4, atol=1e-5)
        with torch.no_grad():
            torch.testing.assert_close(victim.conv_bias_act(x, w, b, 1, 1, 1, 1, act), yb, rtol=2e-6, atol=2e-6)
    model = victim.get_victim_model("efficientdet-d0", device="cuda", image_size=128, seed=5)
    img = torch.rand(2, 128, 128, 3, device="cuda") * 2 - 1
    with torch.no_grad():
        cls_f, box_f = model(img)
        victim.FUSED_EPILOGUE = False
        try:
            cls_t, box_t = model(img)
        finally:
            victim.FUSED_EPILOGUE = True
    for a, b_ in zip(cls_f + box_f, cls_t + box_t):
        torch.testing.assert_close(a, b_, rtol=1e-3, atol=1e-3)


def test_fused_squeeze_excite_matches_torch_ops():
    torch.manual_seed(1)
    blk = victim.MBConv(24, 24, 6, 3, 1).cuda().eval().to(memory_format=torch.channels_last)
    for p in blk.parameters():
        p.requires_grad_(False)
    y = torch.randn(3, 144, 20, 28, device="cuda").contiguous(memory_format=torch.channels_last)
    ya = y.clone().requires_grad_(True)
    oa = victim.squeeze_excite(ya, blk.se_reduce, blk.se_expand)
    yb = y.clone().requires_grad_(True)
    s = yb.mean((2, 3), keepdim=True)
    ob = yb * torch.sigmoid(blk.se_expand(torch.nn.functional.silu(blk.se_reduce(s))))
    g = torch.randn_like(ob)
    oa.backward(g)
    ob.backward(g)
    torch.testing.assert_close(oa, ob, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ya.grad, yb.grad, rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        torch.testing.assert_close(victim.squeeze_excite(y.clone(), blk.se_reduce, blk.se_expand), ob, rtol=1e-5, atol=1e-6)
    # 810-channel identity epilogue (even, not a multiple of 4) with gradient
    x = torch.randn(2, 64, 9, 7, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(810, 64, 1, 1, device="cuda") * 0.1
    b = torch.randn(810, device="cuda")
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    oa = victim.conv_bias_act(xa, w, b, 1, 0, 1, 1, False)
    ob = torch.nn.functional.conv2d(xb, w, b)
    gg = torch.randn_like(ob)
    oa.backward(gg)
    ob.backward(gg)
    torch.testing.assert_close(oa, ob, rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(xa.grad, xb.grad, rtol=1e-4, atol=1e-5)


def test_fused_bifpn_fusion_matches_torch_ops():
    torch.manual_seed(2)
    for n in (2, 3):
        xs = [torch.randn(2, 64, 16, 12, device="cuda").contiguous(memory_format=torch.channels_last) for _ in range(n)]
        w = torch.rand(n, device="cuda")
        w = w / w.sum()
        xa = [x.clone().requires_grad_(True) for x in xs]
        xb = [x.clone().requires_grad_(True) for x in xs]
        xa[1].requires_grad_(False)                                     # an input that needs no gradient is skipped
        oa = victim.fuse_silu(xa, w)
        y = xb[0] * w[0]
        for i in range(1, n):
            y = y + xb[i] * w[i]
        ob = torch.nn.functional.silu(y)
        g = torch.randn_like(ob)
        oa.backward(g)
        ob.backward(g)
        torch.testing.assert_close(oa, ob, rtol=1e-5, atol=1e-6)
        for i in range(n):
            if i == 1:
                assert xa[i].grad is None
            else:
                torch.testing.assert_close(xa[i].grad, xb[i].grad, rtol=1e-4, atol=1e-6)
        with torch.no_grad():
            torch.testing.assert_close(victim.fuse_silu(xs, w), ob, rtol=1e-5, atol=1e-6)


def test_fit_loop_validation_return_values_and_checkpoint(tmp_path):
    """the Keras-facing surface attacker_train.py drives: fit() -> train_step / test_step, epoch means, the
    ModelCheckpoint-style save under `patch_{epoch:02d}_{val_asr_to_scale:.4f}`; call(training=False) returns the second
    pass' (boxes, scores) as the reference does (attacker.py:219) and records the reference's metric names."""
    import os
    from mladversarialobjectdetection_b200 import patch_io
    H, P, B = 128, 24, 2
    torch.manual_seed(0)
    model = victim.get_victim_model("efficientdet-d0", device="cuda", image_size=H)
    att = PatchAttacker(model, patch_size=P, device="cuda", seed=2)
    att.compile(learning_rate=1e-2)
    bt = synth.make_batch(B, H, H, seed=5, max_boxes=2)
    images = torch.from_numpy(bt.images).cuda()
    hist = att.fit([images, images], validation_data=[images], epochs=2, steps_per_epoch=2, validation_steps=1,
                   save_dir=str(tmp_path), verbose=False)
    assert len(hist) == 2
    for k in ("loss", "scale", "scale_loss", "tv_loss", "mean_max_score", "std_max_score", "val_loss", "val_scale_loss",
              "val_tv_loss", "val_asr", "val_asr_to_scale"):
        assert k in hist[-1], k
    saved = sorted(os.listdir(tmp_path))
    assert len(saved) == 2 and saved[0].startswith("patch_01_") and saved[1].startswith("patch_02_")
    p2, s2 = patch_io.load_weights(os.path.join(str(tmp_path), saved[1]))
    np.testing.assert_array_equal(p2, att._patch.cpu().numpy())
    boxes_pred, scores_pred = att(images, training=False)
    assert boxes_pred.nrows() == B and len(scores_pred) == B
    # validation loss = sum(M^2 + (M - scale)^2) + 1e-5 TV, from the same pieces the training step reports
    m = att.metrics
    assert abs(float(m["loss"]) - (float(m["scale_loss"]) + B * (float(m["mean_max_score"]) ** 2 + float(m["std_max_score"]) ** 2)
                                   + 1e-5 * float(m["tv_loss"]))) < 1e-4 * max(1.0, float(m["loss"]))
