"""Host-side mirror of the reference's `adv_patch.py`: `AdversarialPatch`, the inference-time patcher
(adv_patch.py:16-190), with the frame and the patch resident on the GPU.

Same constructor keywords and the same `add_adv_to_img(img, bboxes)` call; every per-pixel step (print adjust, 8-bit
YUV brightness match against the letter-boxed frame, INTER_AREA down-sampling, noise, re-quantisation, paste) runs in
libeotpatch.so (csrc/adv_u8.cu) and is bit-identical to the reference's NumPy + OpenCV code.  The noise of
`random_noise` is drawn on the device (torch, float64) unless given explicitly (parity tests replay NumPy's draws).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream


class AdversarialPatch:
    """add adversarial patch to an image (adv_patch.py:16)"""

    def __init__(self, *, scale, h=640, w=640, patch_file=None, patch=None, device="cuda", seed: Optional[int] = None):
        self.device = torch.device(device)
        if patch is not None:                                      # decoded RGB uint8 array (what Image.open(...) yields)
            raw = np.asarray(patch, dtype=np.uint8)
        elif patch_file is not None:
            from PIL import Image
            raw = np.asarray(Image.open(patch_file).convert("RGB"))
        else:
            raw = (np.random.rand(h, w, 3) * 255).astype("uint8")
        if raw.ndim != 3 or raw.shape[2] != 3:
            raise ValueError("patch must be [h,w,3] RGB")
        self.scale = float(scale)
        self.mean_rgb = 127.
        self.stddev_rgb = 128.
        self.output_size = int(h), int(w)
        self._gen = torch.Generator(device=self.device)
        if seed is not None:
            self._gen.manual_seed(int(seed))
        self._patch_raw = torch.from_numpy(np.ascontiguousarray(raw)).to(self.device)
        self._patch_img = self.print_patch()
        n = ctypes.c_size_t(0)
        _lib.check(_lib.load().adv_u8_workspace_bytes(raw.shape[0], raw.shape[1], ctypes.byref(n)), "adv_u8_workspace_bytes")
        self._ws = torch.empty(int(n.value), dtype=torch.uint8, device=self.device)

    def print_patch(self) -> torch.Tensor:
        """simulate printing and re-imaging with deterministic values (adv_patch.py:40-59)"""
        out = torch.empty_like(self._patch_raw)
        _lib.check(_lib.load().adv_u8_print_patch(_ptr(self._patch_raw), _ptr(out), self._patch_raw.numel(), _stream()),
                   "adv_u8_print_patch")
        return out

    def _create(self, img, bbox):
        """patch coordinates from the person bounding box (adv_patch.py:61-92) -> [ymin, xmin, patch_h, patch_w]"""
        return self.placements(img.shape[0], img.shape[1], [bbox])[0].tolist()

    def placements(self, frame_h: int, frame_w: int, bboxes: Sequence) -> np.ndarray:
        boxes = np.ascontiguousarray(np.asarray(bboxes, dtype=np.float64).reshape(-1, 4))
        out = np.zeros((len(boxes), 4), np.int32)
        _lib.check(_lib.load().adv_u8_box_geometry(int(frame_h), int(frame_w), self.scale,
                                                   boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(boxes),
                                                   out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))), "adv_u8_box_geometry")
        return out

    def add_adv_to_img(self, img, bboxes, noise: Optional[Sequence] = None):
        """add patches to all the persons of one frame (adv_patch.py:179-190).  img: uint8 [h,w,3], NumPy (returns
        NumPy, like the reference) or a CUDA tensor (returns a new CUDA tensor)."""
        was_numpy = not isinstance(img, torch.Tensor)
        frame = (torch.from_numpy(np.ascontiguousarray(img)) if was_numpy else img).to(self.device).clone().contiguous()
        if frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3:
            raise TypeError("img must be uint8 [h,w,3]")
        boxes = np.ascontiguousarray(np.asarray(bboxes, dtype=np.float64).reshape(-1, 4))
        n = len(boxes)
        if n:
            pl = self.placements(frame.shape[0], frame.shape[1], boxes)
            sizes = [int(p[2]) * int(p[3]) * 3 for p in pl]
            if noise is None:
                nz = torch.rand(max(sum(sizes), 1), dtype=torch.float64, device=self.device, generator=self._gen) * 0.02 - 0.01
            else:
                nz = torch.cat([torch.as_tensor(np.asarray(a, np.float64)).reshape(-1) for a in noise]).to(self.device)
                if nz.numel() != sum(sizes):
                    raise ValueError("noise must hold one [patch_h, patch_w, 3] float64 array per box")
            ph, pw = int(self._patch_img.shape[0]), int(self._patch_img.shape[1])
            _lib.check(_lib.load().adv_u8_add_patches(_ptr(frame), int(frame.shape[0]), int(frame.shape[1]), _ptr(self._patch_img),
                                                      ph, pw, self.output_size[0], self.output_size[1], self.scale,
                                                      boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n, _ptr(nz),
                                                      _ptr(self._ws), ctypes.c_size_t(self._ws.numel()), _stream()),
                       "adv_u8_add_patches")
        return frame.cpu().numpy() if was_numpy else frame
