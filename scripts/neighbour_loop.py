"""Runs the kernels of the rows either side of the attack step once each (ncu captures): device NMS, letter-box,
augmentation, uint8 inference twin, and the victim stand-in's epilogue kernels.

    python scripts/neighbour_loop.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mladversarialobjectdetection_b200 import anchors as am, ops
from mladversarialobjectdetection_b200.adv_patch import AdversarialPatch
from mladversarialobjectdetection_b200.anchors import feature_sizes

dev, B, H = "cuda", 64, 512
rng = np.random.default_rng(5)
anc = torch.from_numpy(am.anchor_table((H, H))).to(dev)
A = anc.shape[0]
box = [torch.randn(B, h, w, 36, device=dev) * 0.3 for h, w in feature_sizes((H, H), 7)[3:]]
cand = torch.full((B, A), -1.0, device=dev)
for b in range(B):
    for a0 in rng.integers(0, A - 80, 6):
        idx = torch.from_numpy(a0 + rng.choice(80, 40, replace=False)).to(dev)
        cand[b, idx] = torch.from_numpy(rng.uniform(0.5, 0.99, 40).astype(np.float32)).to(dev)
frames = [torch.from_numpy(rng.integers(0, 256, size=(480, 640, 3), dtype=np.uint8)).to(dev) for _ in range(B)]
ap = AdversarialPatch(scale=0.5, h=640, w=640, patch=rng.integers(0, 256, size=(640, 640, 3), dtype=np.uint8), seed=1)
vframe = frames[0]
x = torch.randn(B, 96, 128, 128, device=dev).contiguous(memory_format=torch.channels_last)
bias = torch.randn(96, device=dev)
gate = torch.rand(B, 96, device=dev)
for it in range(2):
    ops.person_nms(cand, box, anc, (H, H))
    out, sums = ops.letterbox_normalize(frames, (H, H), 127.0, 128.0)
    ops.augment_batch(out, torch.ones(B, dtype=torch.uint8, device=dev), 1.1, 0.05, sums=sums)
    ap.add_adv_to_img(vframe, [(40, 60, 440, 200), (100, 250, 420, 400)])
    y = ops.nhwc_bias_act(x, bias, True, out=torch.empty_like(x))
    ops.nhwc_bias_silu_backward(x, bias, y)
    ops.nhwc_channel_scale(x, gate)
    ops.nhwc_channel_dot(x, y)
    ops.nhwc_fuse_silu([x, y], torch.tensor([0.4, 0.6], device=dev))
torch.cuda.synchronize()
print("done")
