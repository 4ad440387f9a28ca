"""Host-side anchor table for the score kernel (same arithmetic as the reference's
automl/efficientdet/tf2/anchors.py:117-165 and utils.py:509-526; generated once per image size)."""
from __future__ import annotations

from functools import lru_cache
from typing import Sequence, Tuple

import numpy as np


def feature_sizes(image_size: Tuple[int, int], max_level: int):
    sizes = [tuple(image_size)]
    for _ in range(max_level):
        h, w = sizes[-1]
        sizes.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
    return sizes


@lru_cache(maxsize=16)
def anchor_table(image_size: Tuple[int, int], min_level: int = 3, max_level: int = 7, num_scales: int = 3,
                 aspect_ratios: Sequence[float] = (1.0, 2.0, 0.5), anchor_scale: float = 4.0) -> np.ndarray:
    """[A,4] float32 (ymin,xmin,ymax,xmax), ordered level -> y -> x -> (octave-major, aspect-minor)."""
    H, W = image_size
    fs = feature_sizes(image_size, max_level)
    octaves = np.arange(num_scales, dtype=np.float64) / num_scales
    ax = np.sqrt(np.asarray(aspect_ratios, dtype=np.float64))
    ay = 1.0 / ax
    out = []
    for level in range(min_level, max_level + 1):
        stride_y = fs[0][0] / float(fs[level][0])
        stride_x = fs[0][1] / float(fs[level][1])
        half_x = (anchor_scale * stride_x * 2.0 ** octaves)[:, None] * ax[None, :] / 2.0      # [octave, aspect]
        half_y = (anchor_scale * stride_y * 2.0 ** octaves)[:, None] * ay[None, :] / 2.0
        yc = np.arange(stride_y / 2, H, stride_y)
        xc = np.arange(stride_x / 2, W, stride_x)
        yy = yc[:, None, None].repeat(len(xc), 1)                                            # [y, x, 1]
        xx = xc[None, :, None].repeat(len(yc), 0)
        hy = half_y.reshape(1, 1, -1)
        hx = half_x.reshape(1, 1, -1)
        boxes = np.stack([yy - hy, xx - hx, yy + hy, xx + hx], axis=-1)                      # [y, x, 9, 4]
        out.append(boxes.reshape(-1, 4))
    return np.concatenate(out, 0).astype(np.float32)
