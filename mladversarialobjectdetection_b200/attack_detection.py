"""Host-side mirror of the defender's data path: `Masker` (reference: attack_detection.py:321-498).

Forward only -- no gradient crosses the Masker in the reference (attack_detection.py:178-206).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .ragged import RaggedBoxes
from .sampler import TransformSampler, _mix, _unit

CROP = 240          # attack_detection.py:489


class Masker:
    """add patches to image. during training add self-supervised patches, during evaluation the adversarial patch."""

    def __init__(self, patch: torch.Tensor, scale_regressor: torch.Tensor, *args, min_patch_area=4, name=None,
                 seed: int = 0, **kwargs):
        self._patch = patch
        self._scale = scale_regressor
        self.min_patch_area = min_patch_area
        self.name = name
        self.is_training = False
        self.sampler = TransformSampler(seed)
        self.first_image = None       # None: rank * local batch under torch.distributed, else 0 (see attacker.resolve_first_image)
        self._step = 0
        self._workspace = None

    def _train_patches(self, images: torch.Tensor) -> torch.Tensor:
        """tf.random.shuffle(images[:, :240, :240, :]) + random LR / UD flips (attack_detection.py:489-491)."""
        B = images.shape[0]
        g = torch.Generator(device=images.device)
        g.manual_seed(self.sampler.seed * 7919 + self._step)
        perm = torch.randperm(B, device=images.device, generator=g)
        crops = images[perm, :CROP, :CROP, :]
        flip = torch.rand((2, B), device=images.device, generator=g) < 0.5
        crops = torch.where(flip[0].view(B, 1, 1, 1), crops.flip(2), crops)
        crops = torch.where(flip[1].view(B, 1, 1, 1), crops.flip(1), crops)
        return crops.contiguous()

    def __call__(self, inputs, training=False, transforms=None, patches: Optional[torch.Tensor] = None):
        return self.call(inputs, training=training, transforms=transforms, patches=patches)

    def call(self, inputs, training=False, transforms=None, patches: Optional[torch.Tensor] = None):
        """(images', masks) -- attack_detection.py:478-498.  Boxes whose patch area is <= min_patch_area are
        skipped (the reference would index past its filtered list there, attack_detection.py:384)."""
        boxes, images = inputs
        if not isinstance(boxes, RaggedBoxes):
            boxes = RaggedBoxes.from_rows(boxes, images.device)
        self.is_training = training
        n = int(boxes.values.shape[0])
        if training:
            geom = ops.PatchGeometry(tolerance=0.5, noise_amp=0.1, min_patch_area=float(self.min_patch_area), max_scale=0.5)
            patch = patches if patches is not None else self._train_patches(images)
            scale_range = (0.3, 0.5)
        else:
            geom = ops.PatchGeometry(tolerance=0.0, noise_amp=0.1, min_patch_area=float(self.min_patch_area))
            patch = self._patch
            scale_range = None
        if transforms is None:
            from .attacker import resolve_first_image
            params, print_wb = self.sampler.draw(self._step, resolve_first_image(self.first_image, images.shape[0]),
                                                 boxes.row_splits, n, scale_range=scale_range)
            self._step += 1           # (like Patcher: explicit transforms replay a draw, they do not consume one)
        else:
            params, print_wb = transforms
        out, mask, ctx = ops.apply_forward(patch, self._scale, images, boxes.values, boxes.row_splits, params, print_wb,
                                           geom, want_mask=True, workspace=self._workspace)
        self._workspace = ctx.workspace
        return out, mask
