"""Draw of the transform seeds the reference takes from TF's RNG inside `Patcher`
(attacker.py:370-371, 426-427, 436, 473-474; Masker: attack_detection.py:350-351, 411, 421, 451-453).

Counter-based: every number is a hash of (seed, step, GLOBAL image index, box index in image, slot), so a
batch sharded over G ranks draws exactly what the single-GPU batch draws, with no host sync and no state.

On the device the whole draw is ONE kernel of libeotpatch (`eot_draw_transforms`, csrc/eot_draw.cu): `draw`.
`box_params` / `print_wb` below are the same arithmetic in torch ops -- the definition the kernel is tested
against, and what the CPU-only (gloo) tests of the sharding logic use; CUDA callers are routed to the kernel.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

_M1 = -7046029254386353131      # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687      # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685      # 0x94D049BB133111EB


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrap-around arithmetic)."""
    x = x + _M1
    x = (x ^ _lsr(x, 30)) * _M2
    x = (x ^ _lsr(x, 27)) * _M3
    return x ^ _lsr(x, 31)


def _unit(h: torch.Tensor) -> torch.Tensor:
    """top 24 bits -> float32 in [0,1)."""
    return _lsr(h, 40).to(torch.float32) * (1.0 / 16777216.0)


class TransformSampler:
    def __init__(self, seed: int = 0, max_angle: float = 20.0 * math.pi / 180.0, max_delta: float = 0.3,
                 perspective: float = 0.0):
        self.seed = int(seed)
        self.max_angle = max_angle
        self.max_delta = max_delta
        self.perspective = perspective

    def draw(self, step: int, first_image: int, offsets: torch.Tensor, box_capacity: int,
             scale_range: Optional[Tuple[float, float]] = None):
        """(params uint8 [box_capacity,48], print_wb [B,6]) in one kernel launch; the box count in use is read on the
        device (offsets[B]), slots past it are zero."""
        from . import ops
        return ops.draw_transforms(self.seed, step, first_image, offsets, box_capacity, max_angle=self.max_angle,
                                   max_delta=self.max_delta, perspective=self.perspective, scale_range=scale_range)

    def _base(self, step: int, idx: torch.Tensor, slot: int) -> torch.Tensor:
        k = torch.full_like(idx, (self.seed * 1000003 + step) & 0x7FFFFFFFFFFFFFFF)
        return _mix(_mix(k ^ (idx * 0x632BE5AB)) + slot)

    def print_wb(self, step: int, first_image: int, batch: int, device) -> torch.Tensor:
        """[B,6]: w ~ N(.5,.1)^3, b ~ N(0,.01)^3 per image (Box-Muller on hashed uniforms)."""
        img = torch.arange(first_image, first_image + batch, device=device, dtype=torch.int64)
        cols = []
        for c in range(6):
            u1 = _unit(self._base(step, img, 10 + 2 * c)).clamp_min(1e-7)
            u2 = _unit(self._base(step, img, 11 + 2 * c))
            z = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)
            cols.append(0.5 + 0.1 * z if c < 3 else 0.01 * z)
        return torch.stack(cols, dim=1).contiguous()

    def box_params(self, step: int, first_image: int, offsets: torch.Tensor, n_boxes: int,
                   scale_range: Optional[Tuple[float, float]] = None) -> torch.Tensor:
        """uint8 [N,48] EotBoxParams records for the CSR boxes described by `offsets` (int32 [B+1])."""
        device = offsets.device
        buf = torch.zeros((n_boxes, 12), dtype=torch.float32, device=device)
        if n_boxes == 0:
            return buf.view(torch.uint8).reshape(0, 48)
        B = offsets.numel() - 1
        j = torch.arange(n_boxes, device=device, dtype=torch.int64)
        img_local = torch.searchsorted(offsets[1:].to(torch.int64).contiguous(), j, right=True)
        in_img = j - offsets.to(torch.int64)[img_local]
        idx = (img_local + first_image) * 4099 + in_img          # unique per (global image, box in image)
        u = lambda slot: _unit(self._base(step, idx, slot))
        buf[:, 0] = u(40)
        buf[:, 1] = u(41)
        buf[:, 2] = u(42) * (2 * self.max_delta) - self.max_delta
        ang = u(43) * (2 * self.max_angle) - self.max_angle
        buf[:, 3] = torch.cos(ang)
        buf[:, 4] = torch.sin(ang)
        if self.perspective > 0:
            buf[:, 5] = u(44) * (2 * self.perspective) - self.perspective
            buf[:, 6] = u(45) * (2 * self.perspective) - self.perspective
        if scale_range is None:
            buf[:, 7] = -1.0
        else:
            buf[:, 7] = u(46) * (scale_range[1] - scale_range[0]) + scale_range[0]
        keys = self._base(step, idx, 47)
        ib = buf.view(torch.int32)
        ib[:, 8] = (keys & 0xFFFFFFFF).to(torch.int32)           # wraps to the same 32 bits
        ib[:, 9] = (_lsr(keys, 32) & 0xFFFFFFFF).to(torch.int32)
        return buf.view(torch.uint8).reshape(n_boxes, 48)
