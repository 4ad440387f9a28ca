"""CPU port of one `PatchAttacker.train_step` (attacker.py:172-219, 307-316) built from the oracle pieces.

TEST INFRASTRUCTURE / CPU BASELINE (see oracle/__init__.py): used by tests and by bench.py's `cpu_baseline`
and `--impl reference` legs only.  The victim is whatever torch module the caller hands over (moved to CPU);
the patcher, the objective, their gradients and the Adam update are the float32 NumPy restatement -- the same
loops as the reference (per image, per box).
"""
from __future__ import annotations

import numpy as np
import torch

from . import objective, patcher, tfops
from .tfops import F


class AdamState:
    """Keras Adam (ResourceApplyAdam formulas) + clip constraint, float32."""

    def __init__(self, shape, lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-7):
        self.m = np.zeros(shape, F)
        self.v = np.zeros(shape, F)
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0

    def tick(self):
        self.t += 1
        return F(self.lr * np.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t))

    def apply(self, var, grad, alpha, lo, hi):
        self.m = (self.m + (grad - self.m) * F(1 - self.b1)).astype(F)
        self.v = (self.v + (grad * grad - self.v) * F(1 - self.b2)).astype(F)
        return np.clip(var - (self.m * alpha) / (np.sqrt(self.v) + F(self.eps)), F(lo), F(hi)).astype(F)


def victim_heads(victim, images: np.ndarray, need_grad: bool):
    x = torch.from_numpy(np.ascontiguousarray(images))
    if need_grad:
        x.requires_grad_(True)
    cls, box = victim(x, pre_mode=None, post_mode=None)
    return x, cls, box


def attack_step(victim, patch: np.ndarray, scale: float, images: np.ndarray, boxes, params, print_wb,
                anchors: np.ndarray, adam_patch: AdamState = None, adam_scale: AdamState = None,
                first_pass: bool = True, **patcher_kw):
    """One step; returns dict(grad_patch, dscale, loss, max_scores, patched, new_patch, new_scale)."""
    B, H, W, _ = images.shape
    num_classes = victim.config.num_classes
    if first_pass:      # clean pass (attacker.py:180): same victim + pre_nms work, detections replaced by `boxes`
        with torch.no_grad():
            _, cls0, box0 = victim_heads(victim, images, False)
        c0, b0 = objective.merge_levels([c.numpy() for c in cls0], [b.numpy() for b in box0], num_classes)
        objective.second_pass_post(c0, b0, anchors, H, W)
    patched, _, states = patcher.patcher_forward(patch, images, boxes, params, print_wb, scale, **patcher_kw)
    x, cls, box = victim_heads(victim, patched, True)
    cls_np = [c.detach().numpy() for c in cls]
    c_all, b_all = objective.merge_levels(cls_np, [b.detach().numpy() for b in box], num_classes)
    post = objective.objective_forward(c_all, b_all, anchors, H, W, scale)
    dcls_all, dscale = objective.objective_backward(c_all, post, scale)
    dcls = objective.split_levels(dcls_all, [c.shape for c in cls_np])
    torch.autograd.backward(cls, [torch.from_numpy(np.ascontiguousarray(d)) for d in dcls])
    G = x.grad.numpy()
    g_patch = patcher.patcher_backward(G, patch, print_wb, states)
    loss, tv, scale_loss, tv_g = objective.attack_loss(post["max_scores"], scale, patch)
    g_patch = (g_patch + F(1e-5) * tv_g).astype(F)
    out = dict(grad_patch=g_patch, dscale=F(dscale), loss=loss, tv=tv, max_scores=post["max_scores"], patched=patched,
               grad_images=G)
    if adam_patch is not None:
        alpha = adam_patch.tick()
        adam_scale.tick()
        out["new_patch"] = adam_patch.apply(patch, g_patch, alpha, -1.0, 1.0)
        out["new_scale"] = float(adam_scale.apply(np.array([scale], F), np.array([dscale], F), alpha, 0.0, 1.0)[0])
    return out
