"""Child process of tests/test_gpu_variants.py: runs the patch-apply forward + backward on seeded batches with whatever
EOT_* switches the environment holds and writes the outputs (and the library's launch count of one forward call) to
an .npz file.  Environment switches are read once per process by libeotpatch, hence the separate process."""
import sys

import numpy as np
import torch

from mladversarialobjectdetection_b200 import _lib, ops, synth

out_path = sys.argv[1]
lib = _lib.load()
res = {}
for tag, (B, H, P, persp, max_boxes) in {"affine": (6, 256, 64, 0.0, 8), "perspective": (4, 256, 100, 2e-4, 5),
                                         "large_patch": (2, 512, 240, 0.0, 6)}.items():
    bt = synth.make_batch(B, H, H, seed=77, max_boxes=max_boxes, perspective=persp)
    dev = "cuda"
    images = torch.from_numpy(bt.images).to(dev)
    boxes, offsets = torch.from_numpy(bt.boxes).to(dev), torch.from_numpy(bt.offsets).to(dev)
    params, wb = ops.params_to_tensor(bt.params, dev), torch.from_numpy(bt.print_wb).to(dev)
    patch = torch.from_numpy(synth.make_patch(P)).to(dev)
    scale = torch.tensor(0.4, device=dev)
    out = torch.empty_like(images)
    _, _, ctx = ops.apply_forward(patch, scale, images, boxes, offsets, params, wb, out=out)
    torch.cuda.synchronize()
    n0 = lib.eot_launch_count()
    _, _, ctx = ops.apply_forward(patch, scale, images, boxes, offsets, params, wb, out=out, workspace=ctx.workspace)
    torch.cuda.synchronize()
    res[tag + "_launches"] = np.int64(lib.eot_launch_count() - n0)
    G = torch.from_numpy(np.random.default_rng(5).standard_normal(bt.images.shape).astype(np.float32)).to(dev)
    gp = ops.apply_backward(ctx, G)
    torch.cuda.synchronize()
    ops.check_workspace(ctx)                                        # raises on any device-side error flag (lost signal, ...)
    res[tag + "_out"] = out.cpu().numpy()
    res[tag + "_grad"] = gp.cpu().numpy()
np.savez(out_path, **res)
print("ok")
