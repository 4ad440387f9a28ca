"""Regenerates the committed fixtures.  Run in the BUILD container (needs /root/reference):

    python tests/golden/make_golden.py

(a) `adv_patch_create.npz` -- outputs of the REFERENCE's own `adv_patch.AdversarialPatch._create`
    (/root/reference/adv_patch.py:61-92), the only importable piece of the patch-placement logic.
(c) `nms_np.npz`           -- hard and Gaussian soft NMS selections of the REFERENCE's own NumPy implementation
    (/root/reference/automl/efficientdet/nms_np.py:89-191) on seeded boxes.  nms_np uses the pixel-inclusive area
    convention ((x2-x1+1)*(y2-y1+1)) while tf.raw_ops.NonMaxSuppressionV5 does not; the boxes are a few thousand
    units wide so the two IoUs differ by ~1e-3, and cases with an IoU / score that close to a threshold are
    rejected, so that both definitions must select the same boxes in the same order.
(d) `map_fn.npz`           -- outputs of the REFERENCE's own `DataSequence._map_fn`
    (/root/reference/train_data_generator.py:55-75; NumPy + cv2) on seeded uint8 frames.  The module imports
    TensorFlow at the top (absent here); the import is satisfied by an inert stub so that the reference's own
    function body runs unmodified.
(e) `adv_patch_u8.npz`     -- outputs of the REFERENCE's own `adv_patch.AdversarialPatch.add_adv_to_img`
    (/root/reference/adv_patch.py:179-190; NumPy + cv2) on seeded uint8 frames, with the raw patch and the
    np.random.uniform noise draws it consumed (replayed from the same RandomState) stored next to them.
(f) `patcher_ref.npz`      -- output of the REFERENCE's own `attacker.Patcher.call` (/root/reference/attacker.py:344-498,
    with /root/reference/brightness_matcher.py) executed verbatim on a NumPy stand-in for the TF / TFA ops it calls
    (tests/golden/tf_numpy_shim.py: the reference's Python -- control flow, expression order, casts, pad / where / clip /
    scatter sequence, box filter -- is real; the leaf kernels inside the TF wheels are the oracle's restatements).
(g) `objective_ref.npz`    -- outputs of the REFERENCE's own `PatchAttacker.second_pass` / `first_pass` + the max-score line of
    `call` (/root/reference/attacker.py:69-170,190), with the vendored automl code they call imported for real
    (tf2/postprocess.py pre_nms / nms / clip_boxes, tf2/anchors.py, hparams_config.py), all on the NumPy TF shim;
    the NonMaxSuppressionV5 / sigmoid / exp leaf kernels are the oracle's restatements.
(b) `oracle_small.npz`     -- the oracle's forward/backward on a small seeded case, so that the
    oracle cannot drift silently and the GPU box (which has no /root/reference) can check both the
    oracle and the CUDA path against a committed vector.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def gen_adv_patch_create():
    sys.path.insert(0, "/root/reference")
    import adv_patch  # noqa: the reference module itself (NumPy + cv2 + PIL only)
    rng = np.random.default_rng(2022)
    rows = []
    for frame_hw, scale in [((480, 640), 0.5), ((640, 640), 0.4), ((1024, 1024), 0.25)]:
        ap = adv_patch.AdversarialPatch(scale=scale, h=32, w=32)
        img = np.zeros(frame_hw + (3,), np.uint8)
        boxes = [(50, 125, 400, 200)]
        for _ in range(40):
            h = int(rng.integers(8, frame_hw[0])); w = int(rng.integers(8, frame_hw[1]))
            y = int(rng.integers(0, frame_hw[0] - h + 1)); x = int(rng.integers(0, frame_hw[1] - w + 1))
            boxes.append((y, x, y + h, x + w))
        for bb in boxes:
            rows.append([frame_hw[0], frame_hw[1], scale, *bb, *ap._create(img, bb)])
    np.savez_compressed(os.path.join(HERE, "adv_patch_create.npz"), rows=np.asarray(rows, dtype=np.float64))


def gen_oracle_small():
    from mladversarialobjectdetection_b200 import synth
    from oracle import objective, patcher
    bt = synth.make_batch(2, 64, 64, max_boxes=3, min_boxes=2, seed=314)
    patch = synth.make_patch(16, seed=3)
    bx, pr = bt.ragged()
    out, _, states = patcher.patcher_forward(patch, bt.images, bx, pr, bt.print_wb, 0.4)
    G = np.random.default_rng(15).normal(size=out.shape).astype(np.float32)
    gp = patcher.patcher_backward(G, patch, bt.print_wb, states)
    plans = np.array([[bs.plan.y0, bs.plan.x0, bs.plan.ps, bs.plan.d, bs.plan.pad_lo]
                      for st in states for bs in st.boxes], dtype=np.int32)
    rng = np.random.default_rng(16)
    fs = objective.feat_sizes(64, 7)[3:]
    cls = [rng.normal(-2, 2, (2, h, w, 810)).astype(np.float32) for h, w in fs]
    box = [rng.normal(0, 0.5, (2, h, w, 36)).astype(np.float32) for h, w in fs]
    anchors = objective.anchor_boxes(64)
    c, b = objective.merge_levels(cls, box, 90)
    post = objective.objective_forward(c, b, anchors, 64, 64, 0.4)
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), out=out, grad_patch=gp, plans=plans,
                        G_seed=15, max_scores=post["max_scores"], cand_count=post["cand"].sum(1),
                        anchors_head=anchors[:18], anchors_tail=anchors[-9:])


def gen_nms_np():
    sys.path.insert(0, "/root/reference/automl/efficientdet")
    import nms_np  # noqa: the reference's NumPy NMS
    rng = np.random.default_rng(77)
    cases = {}
    k = 0
    while k < 6:
        n = int(rng.integers(20, 120))
        # clusters of overlapping boxes, coordinates in thousands of units
        centres = rng.uniform(2000, 18000, size=(int(rng.integers(3, 9)), 2))
        c = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 400, size=(n, 2))
        wh = rng.uniform(1500, 5000, size=(n, 2))
        x1y1 = c - wh / 2
        x2y2 = c + wh / 2
        scores = rng.uniform(0.05, 1.0, n)
        dets = np.concatenate([x1y1, x2y2, scores[:, None]], 1).astype(np.float64)   # [x1,y1,x2,y2,score]
        hard = nms_np.hard_nms(dets.copy(), 0.5)
        soft = nms_np.soft_nms(dets.copy(), dict(method="gaussian", sigma=0.5, iou_thresh=None, score_thresh=0.2))
        # margins: no pairwise IoU within 0.02 of the hard threshold, no soft score within 0.01 of the score
        # threshold or of another selected score
        a = dets[:, None, :4]; b = dets[None, :, :4]
        iw = np.maximum(np.minimum(a[..., 2], b[..., 2]) - np.maximum(a[..., 0], b[..., 0]), 0)
        ih = np.maximum(np.minimum(a[..., 3], b[..., 3]) - np.maximum(a[..., 1], b[..., 1]), 0)
        area = (dets[:, 2] - dets[:, 0]) * (dets[:, 3] - dets[:, 1])
        iou = iw * ih / (area[:, None] + area[None, :] - iw * ih)
        ss = np.sort(soft[:, 4])
        if (np.abs(iou - 0.5) < 0.02).any() or (np.abs(ss - 0.2) < 0.01).any() or (np.diff(ss) < 1e-3).any():
            continue
        # ... nor may a REJECTED box come that close to the score threshold: the selection must not move with it
        lo = nms_np.soft_nms(dets.copy(), dict(method="gaussian", sigma=0.5, iou_thresh=None, score_thresh=0.19))
        hi = nms_np.soft_nms(dets.copy(), dict(method="gaussian", sigma=0.5, iou_thresh=None, score_thresh=0.21))
        if lo.shape != soft.shape or hi.shape != soft.shape:
            continue
        cases[f"dets{k}"] = dets
        cases[f"hard{k}"] = hard
        cases[f"soft{k}"] = soft
        k += 1
    np.savez_compressed(os.path.join(HERE, "nms_np.npz"), n_cases=k, **cases)


def gen_map_fn():
    import types
    from unittest import mock
    tf = mock.MagicMock()
    tf.keras.utils.Sequence = type("Sequence", (), {})          # DataSequence's base class must be a real class
    stubs = {"tensorflow": tf, "hparams_config": mock.MagicMock(), "util": mock.MagicMock(), "utils": mock.MagicMock()}
    with mock.patch.dict(sys.modules, stubs):
        sys.path.insert(0, "/root/reference")
        import train_data_generator as tdg  # noqa: the reference module; only _map_fn (NumPy + cv2) is executed
    rng = np.random.default_rng(55)
    out = {}
    sizes = [(48, 64), (37, 53), (80, 60), (64, 64), (128, 128), (200, 31), (9, 150)]
    for k, (h, w) in enumerate(sizes):
        frame = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        for tag, mean, std in (("a", 127.0, 128.0), ("b", [123.675, 116.28, 103.53], [58.395, 57.12, 57.375])):
            ds = tdg.DataSequence("", (64, 64), mean, std, file_list=["unused"])
            out[f"frame{k}"] = frame
            out[f"out{k}{tag}"] = ds._map_fn(frame).astype(np.float32)      # generator: tf.convert_to_tensor(.., float32)
    np.savez_compressed(os.path.join(HERE, "map_fn.npz"), n=len(sizes), **out)


def gen_adv_patch_u8():
    sys.path.insert(0, "/root/reference")
    import adv_patch  # noqa
    out = {}
    P = 96
    cases = [((96, 96), [(10, 20, 90, 60), (30, 50, 70, 95)], 0.5),          # frame == output size
             ((72, 96), [(5, 5, 60, 40), (20, 40, 71, 90), (0, 0, 30, 30)], 0.5),
             ((192, 192), [(0, 40, 192, 120), (60, 60, 180, 150)], 0.5),      # exact 2x rescale; first box: patch side == P
             ((150, 211), [(20, 30, 140, 100), (40, 100, 120, 200), (100, 10, 149, 60)], 0.4),
             # boxes whose patch side exceeds the 96 px texture: the INTER_CUBIC branch (adv_patch.py:158-160); this case
             # runs with IPP switched off -- OpenCV's own bicubic kernel, see oracle/adv_patch_u8.py
             ((400, 420), [(10, 20, 390, 200), (30, 150, 330, 400), (100, 100, 200, 160)], 0.5)]
    import cv2
    for k, (hw, boxes, scale) in enumerate(cases):
        cv2.ipp.setUseIPP(k != 4)
        np.random.seed(100 + k)
        raw = (np.random.rand(P, P, 3) * 255).astype("uint8")               # what __init__ draws for patch_file=None
        np.random.seed(100 + k)
        ap = adv_patch.AdversarialPatch(scale=scale, h=P, w=P)
        frame = np.random.default_rng(200 + k).integers(0, 256, size=hw + (3,), dtype=np.uint8)
        state = np.random.get_state()
        res = ap.add_adv_to_img(frame, boxes)
        np.random.set_state(state)                                           # replay the noise draws of random_noise()
        noises = []
        for bb in boxes:
            _, _, ph, pw = ap._create(frame, bb)
            noises.append(np.random.uniform(low=-0.01, high=0.01, size=(ph, pw, 3)))
        out[f"raw{k}"] = raw
        out[f"printed{k}"] = ap._patch_img
        out[f"frame{k}"] = frame
        out[f"boxes{k}"] = np.asarray(boxes, np.float64)
        out[f"scale{k}"] = scale
        out[f"result{k}"] = res
        for i, nz in enumerate(noises):
            out[f"noise{k}_{i}"] = nz
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(os.path.join(HERE, "adv_patch_u8.npz"), n=len(cases), P=P, **out)


def gen_patcher_ref2():
    """A second, harder Patcher case for the oracle (CPU test only): 3 images of 96x96, 32x32 patch, scale .6 (patch sides
    from 6 to 46 px: down- AND up-sampling), boxes that overlap each other, boxes whose window is clamped at the image
    border (attacker.py:480-486) and one dropped by the area filter."""
    import tf_numpy_shim as shim
    from mladversarialobjectdetection_b200 import synth
    from oracle import patcher, tfops
    F = np.float32
    attacker = shim.import_reference_attacker()
    H = W = 96
    P = 32
    rng = np.random.default_rng(4242)
    images = rng.uniform(-1, 1, (3, H, W, 3)).astype(F)
    patch = synth.make_patch(P, seed=9)
    boxes = [np.array([[5, 5, 80, 40], [20, 25, 90, 60], [0, 60, 30, 95], [60, 0, 95, 20]], F),        # overlapping + corners
             np.array([[10, 10, 20, 18], [2, 3, 6, 5], [30, 30, 94, 94]], F),                            # small, filtered, huge
             np.array([[40, 40, 95, 95], [45, 45, 90, 92]], F)]                                          # nested at the border
    scale = F(0.6)
    params = [np.zeros(len(b), patcher.BOX_PARAMS) for b in boxes]
    print_wb = np.zeros((3, 6), F)
    q = shim.QUEUE
    q.items.clear()
    for b in range(3):
        zw, zb = rng.standard_normal(3).astype(F), rng.standard_normal(3).astype(F)
        print_wb[b, :3] = zw * F(0.1) + F(0.5)
        print_wb[b, 3:] = zb * F(0.01) + F(0.0)
        q.push("normal", zw)
        q.push("normal", zb)
        plans = []
        for j in range(len(boxes[b])):
            params[b][j]["uy"], params[b][j]["ux"] = F(rng.random()), F(rng.random())
            params[b][j]["scale"] = F(-1.0)
            params[b][j]["key0"], params[b][j]["key1"] = int(rng.integers(0, 2 ** 31)), int(rng.integers(0, 2 ** 31))
            q.push("uniform", F(params[b][j]["uy"]))
            q.push("uniform", F(params[b][j]["ux"]))
            plans.append(patcher.create(boxes[b][j], scale, params[b][j]["uy"], params[b][j]["ux"], 0.2, H, W))
        for j, pl in enumerate(plans):
            if not pl.valid:
                continue
            n = pl.ps * pl.ps * 3
            words = tfops.philox4x32_10(np.arange((n + 3) // 4, dtype=np.uint32), int(params[b][j]["key0"]),
                                        int(params[b][j]["key1"])).reshape(-1)[:n]
            q.push("uniform", ((words & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(F) - F(1.0))
            ud, ua = F(rng.random()), F(rng.random())
            params[b][j]["delta"] = ud * (F(0.3) - F(-0.3)) + F(-0.3)
            q.push("uniform", ud)
            lo, hi = F(-20.0 * np.pi / 180.0), F(20.0 * np.pi / 180.0)
            ang = F(ua * (hi - lo) + lo)
            params[b][j]["cos"], params[b][j]["sin"] = F(np.cos(ang)), F(np.sin(ang))
            q.push("uniform", ua)
    layer = attacker.Patcher(shim.Variable(patch.astype(F)), shim.Variable(scale), name="Patcher")
    out_ref = np.asarray(layer([boxes, images]), F)
    assert not q.items
    out_oracle, _, states = patcher.patcher_forward(patch, images, boxes, params, print_wb, float(scale))
    assert np.array_equal(out_ref, out_oracle), "oracle != reference Patcher on the shim (case 2)"
    sizes = sorted(bs.plan.ps for st in states for bs in st.boxes)
    assert sizes[0] < P < sizes[-1], sizes                                   # both resize directions are exercised
    offsets = np.cumsum([0] + [len(b) for b in boxes]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "patcher_ref2.npz"), patch=patch, images=images, boxes=np.concatenate(boxes),
                        offsets=offsets, params=np.concatenate(params).view(np.uint8), print_wb=print_wb, scale=scale,
                        out_ref=out_ref)


def gen_patcher_ref():
    import tf_numpy_shim as shim
    from mladversarialobjectdetection_b200 import synth
    from oracle import patcher, tfops
    F = np.float32
    attacker = shim.import_reference_attacker()
    H = W = 64
    P = 16
    bt = synth.make_batch(2, H, W, max_boxes=3, min_boxes=2, seed=2718)
    patch = synth.make_patch(P, seed=5)
    boxes, params = bt.ragged()
    boxes = [np.array(b, F) for b in boxes]
    params = [np.array(p) for p in params]
    # one box the area filter drops (patch side floor(4 * .4) = 1 -> area 1 <= 4), in the middle of image 0's list
    tiny = np.array([[10.0, 12.0, 14.0, 15.0]], F)
    boxes[0] = np.concatenate([boxes[0][:1], tiny, boxes[0][1:]])
    params[0] = np.concatenate([params[0][:1], params[0][:1], params[0][1:]])
    scale = F(0.4)
    rng = np.random.default_rng(99)
    print_wb = np.zeros((2, 6), F)
    q = shim.QUEUE
    q.items.clear()
    for b in range(2):
        zw, zb = rng.standard_normal(3).astype(F), rng.standard_normal(3).astype(F)
        print_wb[b, :3] = zw * F(0.1) + F(0.5)                                   # tf.random.normal((1,1,3), .5, .1)
        print_wb[b, 3:] = zb * F(0.01) + F(0.0)
        q.push("normal", zw)
        q.push("normal", zb)
        plans = []
        for j in range(len(boxes[b])):                                           # create(): jitter draws of every box
            q.push("uniform", F(params[b][j]["uy"]))
            q.push("uniform", F(params[b][j]["ux"]))
            plans.append(patcher.create(boxes[b][j], scale, params[b][j]["uy"], params[b][j]["ux"], 0.2, H, W))
        for j, pl in enumerate(plans):                                           # the while loop over the VALID boxes
            if not pl.valid:
                continue
            n = pl.ps * pl.ps * 3
            words = tfops.philox4x32_10(np.arange((n + 3) // 4, dtype=np.uint32), int(params[b][j]["key0"]),
                                        int(params[b][j]["key1"])).reshape(-1)[:n]
            q.push("uniform", ((words & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(F) - F(1.0))
            ud, ua = F(rng.random()), F(rng.random())
            lo, hi = F(-0.3), F(0.3)
            params[b][j]["delta"] = ud * (hi - lo) + lo                          # tf.image.random_brightness(im, .3)
            q.push("uniform", ud)
            lo, hi = F(-20.0 * np.pi / 180.0), F(20.0 * np.pi / 180.0)
            ang = F(ua * (hi - lo) + lo)
            params[b][j]["cos"], params[b][j]["sin"] = F(np.cos(ang)), F(np.sin(ang))
            q.push("uniform", ua)
    layer = attacker.Patcher(shim.Variable(patch.astype(F)), shim.Variable(scale), name="Patcher")
    out_ref = np.asarray(layer([boxes, bt.images.astype(F)]), F)
    assert not q.items, "the reference consumed fewer random draws than queued"
    out_oracle, _, _ = patcher.patcher_forward(patch, bt.images, boxes, params, print_wb, float(scale))
    assert np.array_equal(out_ref, out_oracle), "oracle != reference Patcher on the shim"
    offsets = np.cumsum([0] + [len(b) for b in boxes]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "patcher_ref.npz"), patch=patch, images=bt.images, boxes=np.concatenate(boxes),
                        offsets=offsets, params=np.concatenate(params).view(np.uint8), print_wb=print_wb, scale=scale,
                        out_ref=out_ref)
    # the stand-alone BrightnessMatcher layer as well (brightness_matcher.py:43-73)
    bm = attacker.brightness_matcher.BrightnessMatcher(name="bm")
    src = np.random.default_rng(7).uniform(-1, 1, (P, P, 3)).astype(F)
    np.savez_compressed(os.path.join(HERE, "brightness_ref.npz"), src=src, tgt=bt.images[0],
                        out_ref=np.asarray(bm((src, bt.images[0])), F))


def gen_masker_ref():
    """The reference's own `attack_detection.Masker.call` (attack_detection.py:321-498) on the NumPy TF shim, evaluation
    branch (learned patch, tolerance 0) and training branch (per-image crops of other batch images, scale U(.3,.5))."""
    import tf_numpy_shim as shim
    from mladversarialobjectdetection_b200 import synth
    from oracle import patcher, tfops
    F = np.float32
    ad = shim.import_reference_attacker(module="attack_detection")
    H = W = 64
    out = {}
    for tag, training in (("eval", False), ("train", True)):
        bt = synth.make_batch(2, H, W, max_boxes=2, min_boxes=1, seed=161 if training else 162,
                              scale_range=(0.3, 0.5) if training else None)
        boxes, params = bt.ragged()
        boxes = [np.array(b, F) for b in boxes]
        params = [np.array(p) for p in params]
        images = bt.images.astype(F)
        rng = np.random.default_rng(17)
        q = shim.QUEUE
        q.items.clear()
        tol = 0.5 if training else 0.0
        shared = synth.make_patch(24, seed=6)
        scale = F(0.45)
        if training:
            perm = np.array([1, 0])
            fl_lr, fl_ud = np.array([True, False]), np.array([False, True])
            q.push("perm", perm)
            q.push("flip", fl_lr)
            q.push("flip", fl_ud)
            patches = images[:, :240, :240, :][perm].copy()
            patches[fl_lr] = patches[fl_lr][:, :, ::-1]
            patches[fl_ud] = patches[fl_ud][:, ::-1]
        print_wb = np.zeros((2, 6), F)
        for b in range(2):
            plans = []
            for j in range(len(boxes[b])):                                       # create(): [scale], jitter y, jitter x
                if training:
                    us = F(rng.random())
                    params[b][j]["scale"] = us * (F(0.5) - F(0.3)) + F(0.3)
                    q.push("uniform", us)
                else:
                    params[b][j]["scale"] = F(-1.0)
                q.push("uniform", F(params[b][j]["uy"]))
                q.push("uniform", F(params[b][j]["ux"]))
                sc = params[b][j]["scale"] if training else scale
                plans.append(patcher.create(boxes[b][j], sc, params[b][j]["uy"], params[b][j]["ux"], tol, H, W))
            assert all(p.valid for p in plans)
            zw, zb = rng.standard_normal(3).astype(F), rng.standard_normal(3).astype(F)   # print adjust AFTER create here
            print_wb[b, :3] = zw * F(0.1) + F(0.5)
            print_wb[b, 3:] = zb * F(0.01) + F(0.0)
            q.push("normal", zw)
            q.push("normal", zb)
            for j, pl in enumerate(plans):
                n = pl.ps * pl.ps * 3
                words = tfops.philox4x32_10(np.arange((n + 3) // 4, dtype=np.uint32), int(params[b][j]["key0"]),
                                            int(params[b][j]["key1"])).reshape(-1)[:n]
                q.push("uniform", ((words & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(F) - F(1.0))
                ud, ua = F(rng.random()), F(rng.random())
                params[b][j]["delta"] = ud * (F(0.3) - F(-0.3)) + F(-0.3)
                q.push("uniform", ud)
                lo, hi = F(-20.0 * np.pi / 180.0), F(20.0 * np.pi / 180.0)
                ang = F(ua * (hi - lo) + lo)
                params[b][j]["cos"], params[b][j]["sin"] = F(np.cos(ang)), F(np.sin(ang))
                q.push("uniform", ua)
        layer = ad.Masker(shim.Variable(shared.astype(F)), shim.Variable(scale), name="Masker")
        ref_img, ref_mask = layer([boxes, images], training=training)
        assert not q.items
        o_img, o_mask, _ = patcher.patcher_forward(patches if training else shared, images, boxes, params, print_wb, float(scale),
                                                   tolerance=tol, noise_amp=0.1, want_mask=True)
        assert np.array_equal(np.asarray(ref_img, F), o_img) and np.array_equal(np.asarray(ref_mask, F), o_mask), tag
        offsets = np.cumsum([0] + [len(b) for b in boxes]).astype(np.int32)
        out.update({f"{tag}_images": images, f"{tag}_boxes": np.concatenate(boxes), f"{tag}_offsets": offsets,
                    f"{tag}_params": np.concatenate(params).view(np.uint8), f"{tag}_print_wb": print_wb,
                    f"{tag}_patch": patches if training else shared, f"{tag}_scale": scale,
                    f"{tag}_out": np.asarray(ref_img, F), f"{tag}_mask": np.asarray(ref_mask, F)})
    np.savez_compressed(os.path.join(HERE, "masker_ref.npz"), **out)


def gen_anchors_ref():
    """The reference's own `tf2/anchors.py:Anchors` (NumPy arithmetic; TensorFlow and the object_detection imports of the
    module stubbed out) for the image sizes / anchor scales of the configs: SHA-256 of the float32 table + every 97th row."""
    import hashlib
    import importlib.abc
    import importlib.machinery
    from unittest import mock

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def __init__(self):
            self.created = []

        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in ("tensorflow", "object_detection", "absl", "tensorflow_addons", "tensorflow_model_optimization"):
                return importlib.machinery.ModuleSpec(name, self, is_package=True)

        def create_module(self, spec):
            m = mock.MagicMock(name=spec.name)
            m.__path__, m.__name__, m.__spec__, m.__loader__ = [], spec.name, spec, self
            self.created.append(spec.name)
            return m

        def exec_module(self, module):
            pass
    finder = Finder()
    sys.meta_path.insert(0, finder)
    paths = ["/root/reference/automl/efficientdet", "/root/reference/automl/efficientdet/tf2"]
    sys.path[:0] = paths
    saved = {k: sys.modules.pop(k) for k in ("anchors", "utils") if k in sys.modules}
    try:
        import tensorflow as tf
        tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, np.float32)
        import anchors as ref_anchors  # noqa: the reference module
        out = {}
        cfgs = [(512, 512, 4.0), (640, 640, 4.0), (1024, 1024, 4.0), (320, 320, 3.0), (384, 640, 4.0)]
        for k, (h, w, scale) in enumerate(cfgs):
            tab = np.asarray(ref_anchors.Anchors(3, 7, 3, [1.0, 2.0, 0.5], scale, (h, w)).boxes, np.float32)
            out[f"sha{k}"] = hashlib.sha256(np.ascontiguousarray(tab).tobytes()).hexdigest()
            out[f"rows{k}"] = tab[::97]
            out[f"n{k}"] = len(tab)
        np.savez_compressed(os.path.join(HERE, "anchors_ref.npz"), cfgs=np.asarray(cfgs, np.float64), **out)
    finally:
        sys.meta_path.remove(finder)
        for p in paths:
            sys.path.remove(p)
        for name in finder.created + ["anchors", "utils"]:
            sys.modules.pop(name, None)
        sys.modules.update(saved)


def gen_objective_ref():
    import tf_numpy_shim as shim
    from oracle import nms as onms, objective
    from tests._util import objective_fixture_inputs
    F = np.float32
    att = shim.import_reference_attacker(with_automl=True)
    cfg = att.hparams_config.get_efficientdet_config("efficientdet-d0")
    cfg.override({"nms_configs": {"iou_thresh": .5, "score_thresh": .5}})          # attacker_train.py:31
    B, H = 2, 64
    cfg.image_size = H
    cls, box = objective_fixture_inputs(8, B, H)

    class Model:
        config = cfg

        def __call__(self, images, pre_mode=None, post_mode=None):
            return tuple(c.copy() for c in cls), tuple(b.copy() for b in box)
    pa = object.__new__(att.PatchAttacker)
    pa.config, pa.model = cfg, Model()
    images = np.zeros((B, H, H, 3), F)
    boxes_pred, scores_pred = pa.second_pass(images)
    max_scores = np.maximum(shim._reduce_max(scores_pred, axis=1), F(0.0))           # attacker.py:190
    fp_boxes, fp_scores = pa.first_pass(images)
    # cross-check against the oracle right here
    anchors = objective.anchor_boxes(H)
    c_all, b_all = objective.merge_levels(cls, box, 90)
    post = objective.objective_forward(c_all, b_all, anchors, H, H, 0.4)
    for b in range(B):
        assert np.array_equal(post["score"][b][post["cand"][b]], scores_pred.rows[b])
        assert np.array_equal(post["boxes"][b][post["cand"][b]], boxes_pred.rows[b])
    assert np.array_equal(post["max_scores"], max_scores)
    cand = np.where(post["cand"], post["score"], F(-1.0)).astype(F)
    rows, row_scores = onms.person_boxes_after_nms(cand, b_all, anchors, (H, H), dict(cfg.nms_configs.as_dict()), thresh=True)
    for b in range(B):
        assert np.array_equal(rows[b], fp_boxes.rows[b]) and np.array_equal(row_scores[b], fp_scores.rows[b])
    out = dict(seed=8, B=B, H=H, max_scores=max_scores)
    for b in range(B):
        out[f"sp_scores{b}"], out[f"sp_boxes{b}"] = scores_pred.rows[b], boxes_pred.rows[b]
        out[f"fp_scores{b}"], out[f"fp_boxes{b}"] = fp_scores.rows[b], fp_boxes.rows[b]
    # the hard-NMS branch of postprocess.nms (tf2/postprocess.py:175-180) through the same reference code
    cfg.override({"nms_configs": {"method": "hard", "iou_thresh": .5, "score_thresh": .4}})
    hb, hs = pa.first_pass(images)
    hrows, hscores = onms.person_boxes_after_nms(cand, b_all, anchors, (H, H), dict(cfg.nms_configs.as_dict()), thresh=True)
    for b in range(B):
        assert np.array_equal(hrows[b], hb.rows[b]) and np.array_equal(hscores[b], hs.rows[b])
        out[f"hard_scores{b}"], out[f"hard_boxes{b}"] = hs.rows[b], hb.rows[b]
    # attack success rate (attacker.py:238-255) of the reference on ragged score lists: clean pass = the hard-NMS result,
    # "attacked" pass = the same boxes with scores scaled down, swept over PatchAttacker.bins (attacker.py:66)
    rng = np.random.default_rng(12)
    atk = shim.Ragged([(r * rng.uniform(0.3, 1.0, len(r))).astype(F) for r in hs.rows])
    bins = np.arange(0.4, .805, .01, dtype="float32")
    out["asr_bins"] = bins
    out["asr"] = np.asarray([float(pa.calc_asr(hs, atk, hb, hb, score_thresh=float(t))) for t in bins], F)
    for b in range(B):
        out[f"atk_scores{b}"] = atk.rows[b]
    np.savez_compressed(os.path.join(HERE, "objective_ref.npz"), **out)


if __name__ == "__main__":
    gen_adv_patch_create()
    gen_objective_ref()
    gen_anchors_ref()
    gen_masker_ref()
    gen_patcher_ref()
    gen_patcher_ref2()
    gen_adv_patch_u8()
    gen_map_fn()
    gen_nms_np()
    gen_oracle_small()
    print("fixtures written to", HERE)
