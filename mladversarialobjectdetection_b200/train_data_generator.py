"""Host-side mirror of the reference's `train_data_generator.py` for the part that feeds the attack step:
`DataSequence` (train_data_generator.py:24-104) and the augmentation chain of `get_tf_dataset` (:201-226).

File reading / JPEG decoding stays on the host (PIL, as the reference); everything per pixel -- standardise,
aspect-preserving resize, zero pad, flips, contrast, brightness, clip -- runs in libeotpatch.so on the device
(csrc/input_pipeline.cu), one launch per 64 frames instead of a GIL-bound Python generator feeding tf.data.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops


def _read_image(img_dir: str, filename: str) -> np.ndarray:
    """train_data_generator.py:122-132."""
    from PIL import Image
    im = Image.open(os.path.join(img_dir, filename))
    if im.mode != "RGB":
        im = im.convert("RGB")
    return np.asarray(im)


class DataSequence:
    """Sequence subclass to define rescaling and preprocessing operations (train_data_generator.py:24)."""

    def __init__(self, img_dir, output_size, mean_rgb, stddev_rgb, *, file_list=None, shuffle=True, device="cuda"):
        self._img_dir = img_dir
        self._output_size = tuple(int(v) for v in output_size)
        self._mean_rgb = mean_rgb
        self._stddev_rgb = stddev_rgb
        self._flist = list(file_list) if file_list else sorted(os.listdir(self._img_dir))
        self._shuffle = shuffle
        self.device = torch.device(device)

    def __len__(self):
        return len(self._flist)

    def _map_fn(self, image) -> torch.Tensor:
        """preprocessing function (train_data_generator.py:55-75) for one decoded frame -> [H,W,3] float32."""
        return self.map_batch([image])[0][0]

    def map_batch(self, frames: Sequence, out: Optional[torch.Tensor] = None):
        """`_map_fn` over a list of decoded uint8 frames (NumPy or torch, any sizes) -> ([B,H,W,3], channel sums)."""
        dev_frames = [f if isinstance(f, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(f)) for f in frames]
        dev_frames = [f.to(self.device, non_blocking=True) for f in dev_frames]
        return ops.letterbox_normalize(dev_frames, self._output_size, self._mean_rgb, self._stddev_rgb, out=out)

    def __getitem__(self, idx) -> torch.Tensor:
        return self._map_fn(_read_image(self._img_dir, self._flist[idx]))

    def batches(self, batch_size: int, rng: Optional[np.random.Generator] = None):
        """Yields ([B,H,W,3] float32, channel sums) forever, reshuffling per epoch when shuffle=True (:89-104)."""
        rng = rng or np.random.default_rng()
        order = np.arange(len(self))
        while True:
            if self._shuffle:
                rng.shuffle(order)
            for i in range(0, len(order) - batch_size + 1, batch_size):
                yield self.map_batch([_read_image(self._img_dir, self._flist[j]) for j in order[i:i + batch_size]])


def augment(batch: torch.Tensor, rng: np.random.Generator, sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The training-only map chain of get_tf_dataset (train_data_generator.py:218-222): random_flip_left_right, then
    Sequential[RandomFlip('horizontal'), RandomContrast(.2)], random_brightness(.2), clip_by_value(-1, 1)."""
    B = batch.shape[0]
    flip = rng.integers(0, 2, B) ^ rng.integers(0, 2, B)               # two independent coin flips per image
    contrast = float(rng.uniform(0.8, 1.2))
    delta = float(rng.uniform(-0.2, 0.2))
    return ops.augment_batch(batch, torch.from_numpy(flip.astype(np.uint8)).to(batch.device), contrast, delta, sums=sums)
