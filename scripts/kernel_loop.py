"""Runs only this repo's kernels on config-2 shapes (no victim): used for ncu captures and quick timing.

    python scripts/kernel_loop.py [--batch 64] [--image 512] [--patch 100] [--iters 3] [--what fwd,bwd,score]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mladversarialobjectdetection_b200 import anchors as anchors_mod, ops, synth
from mladversarialobjectdetection_b200.anchors import feature_sizes

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--image", type=int, default=512)
ap.add_argument("--patch", type=int, default=100)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--what", default="fwd,bwd,score")
ap.add_argument("--max-boxes", type=int, default=8)
ap.add_argument("--time", action="store_true")
ap.add_argument("--graph", action="store_true", help="time replays of a CUDA graph of the call(s) instead of eager launches")
args = ap.parse_args()
B, H, P = args.batch, args.image, args.patch
dev = "cuda"
bt = synth.make_batch(B, H, H, max_boxes=args.max_boxes)
images = torch.from_numpy(bt.images).to(dev)
boxes, offsets = torch.from_numpy(bt.boxes).to(dev), torch.from_numpy(bt.offsets).to(dev)
params, wb = ops.params_to_tensor(bt.params, dev), torch.from_numpy(bt.print_wb).to(dev)
patch = torch.from_numpy(synth.make_patch(P)).to(dev)
scale = torch.tensor(0.4, device=dev)
out = torch.empty_like(images)
geo = ops.PatchGeometry()
_, _, ctx = ops.apply_forward(patch, scale, images, boxes, offsets, params, wb, geo, out=out)
G = torch.randn_like(images)
gp = torch.empty_like(patch)
what = args.what.split(",")
if any(w.startswith("score") for w in what):
    fs = feature_sizes((H, H), 7)[3:]
    # like the calibrated random-init victim heads: every class equally likely to be the arg-max (person ~ 1/90)
    cls = [torch.randn(B, h, w, 810, device=dev) * 1.5 - 4.6 for h, w in fs]
    box = [torch.randn(B, h, w, 36, device=dev) * 0.3 for h, w in fs]
    anc = torch.from_numpy(anchors_mod.anchor_table((H, H))).to(dev)


state = {}


def one():
    res = {}
    if "fwd" in what:
        state["ctx"] = ops.apply_forward(patch, scale, images, boxes, offsets, params, wb, geo, out=out, workspace=ctx.workspace)[2]
    if "bwd" in what:
        ops.apply_backward(state.get("ctx", ctx), G, grad_patch=gp)
    if "score" in what or "scoref" in what:
        r = ops.score_max_forward(cls, box, anc, (H, H))
        state["sctx"] = r[3]
    if "score" in what or "scoreb" in what:
        ops.score_max_backward(state["sctx"], scale)


if any(w.startswith("score") for w in what):
    state["sctx"] = ops.score_max_forward(cls, box, anc, (H, H))[3]
for _ in range(args.warmup):
    one()
torch.cuda.synchronize()
if args.time:
    for name in what:
        what_saved, what = what, [name]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        one(); torch.cuda.synchronize()
        run = one
        if args.graph:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                one()
            run = gr.replay
            run(); torch.cuda.synchronize()
        e0.record()
        for _ in range(args.iters):
            run()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / args.iters * 1e3:.1f} us per call")
        what = what_saved
else:
    for _ in range(args.iters):
        one()
    torch.cuda.synchronize()
print("done")
