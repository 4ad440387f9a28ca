"""Shared helpers of the GPU parity tests."""
import numpy as np
import torch

from mladversarialobjectdetection_b200 import ops, synth


def to_device(bt: synth.Batch, dev="cuda"):
    return dict(images=torch.from_numpy(bt.images).to(dev),
                boxes=torch.from_numpy(bt.boxes).to(dev),
                offsets=torch.from_numpy(bt.offsets).to(dev),
                params=ops.params_to_tensor(bt.params, dev),
                print_wb=torch.from_numpy(bt.print_wb).to(dev))


def run_forward(patch_np, scale, bt, geom=None, want_mask=False, patch_t=None):
    dev = "cuda"
    d = to_device(bt, dev)
    patch = patch_t if patch_t is not None else torch.from_numpy(patch_np).to(dev)
    sc = torch.tensor(scale, dtype=torch.float32, device=dev)
    geom = geom or ops.PatchGeometry()
    out, mask, ctx = ops.apply_forward(patch, sc, d["images"], d["boxes"], d["offsets"], d["params"], d["print_wb"],
                                       geom, want_mask=want_mask)
    torch.cuda.synchronize()
    return out, mask, ctx, d


def objective_fixture_inputs(seed=8, B=2, H=64):
    """Seeded victim outputs of the objective fixtures (tests/golden/objective_ref.npz); lives in the package
    (synth.objective_inputs) because __graft_entry__.smoke() uses it too."""
    return synth.objective_inputs(seed, B, H)
