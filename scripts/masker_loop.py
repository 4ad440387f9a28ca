"""Config 5 data path: `Masker` forward (attack_detection.py:478-498) at the defender's shapes -- 24 images of 640x640,
per-image 240x240 crops of other batch images as patch textures (strided views, no copy), mask output.

    python scripts/masker_loop.py [--batch 24] [--image 640] [--crop 240] [--iters 50]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mladversarialobjectdetection_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=24)
ap.add_argument("--image", type=int, default=640)
ap.add_argument("--crop", type=int, default=240)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
B, H, P = args.batch, args.image, args.crop
dev = "cuda"
bt = synth.make_batch(B, H, H, max_boxes=8, scale_range=(0.3, 0.5))
images = torch.from_numpy(bt.images).to(dev)
boxes, offsets = torch.from_numpy(bt.boxes).to(dev), torch.from_numpy(bt.offsets).to(dev)
params, wb = ops.params_to_tensor(bt.params, dev), torch.from_numpy(bt.print_wb).to(dev)
geom = ops.PatchGeometry(tolerance=0.5, noise_amp=0.1, max_scale=0.5)
patches = images.roll(1, 0)[:, :P, :P, :]                      # another image's crop, zero-copy strided view
scale = torch.tensor(0.4, device=dev)
out = torch.empty_like(images)
_, _, ctx = ops.apply_forward(patches, scale, images, boxes, offsets, params, wb, geom, want_mask=True, out=out)


def one():
    ops.apply_forward(patches, scale, images, boxes, offsets, params, wb, geom, want_mask=True, out=out,
                      workspace=ctx.workspace)


for _ in range(3):
    one()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    one()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / args.iters * 1e3
alg = 36.0 * H * H * B
print(f"masker forward: {us:.1f} us per call, {B / (us * 1e-6):.0f} images/s, algorithmic {alg / 1e6:.1f} MB "
      f"-> {alg / (us * 1e-6) / 1e9:.0f} GB/s")
