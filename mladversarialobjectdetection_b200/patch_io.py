"""On-disk format of the attack state, as `PatchAttacker.save_weights` writes it and
`PatchAttacker(initial_patch=dir)` / `PatchAttackDefender(eval_patch=...)` read it back
(reference: attacker.py:45-48, 328-341): `scale.txt` (Python literal), `patch.png` (uint8, de-normalised with
stddev/mean) and `patch.tiff` (the raw float32 [-1,1] patch).

tifffile is not available in this image, so the float32 TIFF is written/read by a minimal baseline-TIFF codec
(little-endian, one strip, SampleFormat=IEEE float), which tifffile.imread / imwrite interoperate with.
"""
from __future__ import annotations

import ast
import os
import struct

import numpy as np

_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 11: "f", 12: "d", 16: "Q"}


def write_tiff_f32(path: str, arr: np.ndarray) -> None:
    arr = np.ascontiguousarray(arr, dtype="<f4")
    if arr.ndim == 2:
        arr = arr[..., None]
    h, w, c = arr.shape
    data = arr.tobytes()
    entries = [
        (256, 4, 1, w), (257, 4, 1, h),
        (258, 3, c, None),                 # BitsPerSample (offset filled below when c > 2)
        (259, 3, 1, 1),                    # no compression
        (262, 3, 1, 2 if c >= 3 else 1),   # RGB / BlackIsZero
        (273, 4, 1, None),                 # StripOffsets
        (277, 3, 1, c),                    # SamplesPerPixel
        (278, 4, 1, h),                    # RowsPerStrip
        (279, 4, 1, len(data)),            # StripByteCounts
        (284, 3, 1, 1),                    # PlanarConfiguration: chunky
        (339, 3, c, None),                 # SampleFormat = 3 (IEEE float)
    ]
    n = len(entries)
    ifd_off = 8
    extra_off = ifd_off + 2 + n * 12 + 4
    bits = struct.pack("<%dH" % c, *([32] * c))
    fmt = struct.pack("<%dH" % c, *([3] * c))
    bits_off, fmt_off = extra_off, extra_off + len(bits)
    data_off = (fmt_off + len(fmt) + 15) // 16 * 16
    out = bytearray(struct.pack("<2sHI", b"II", 42, ifd_off))
    out += struct.pack("<H", n)
    for tag, typ, cnt, val in entries:
        if tag == 258:
            val = bits_off if c > 2 else (32 | (32 << 16) if c == 2 else 32)
        elif tag == 339:
            val = fmt_off if c > 2 else (3 | (3 << 16) if c == 2 else 3)
        elif tag == 273:
            val = data_off
        if typ == 3 and cnt == 1:
            out += struct.pack("<HHIHH", tag, typ, cnt, val, 0)
        else:
            out += struct.pack("<HHII", tag, typ, cnt, val)
    out += struct.pack("<I", 0)
    out += bits + fmt
    out += b"\0" * (data_off - len(out))
    out += data
    with open(path, "wb") as f:
        f.write(out)


def read_tiff_f32(path: str) -> np.ndarray:
    raw = open(path, "rb").read()
    bo = {b"II": "<", b"MM": ">"}[raw[:2]]
    magic, ifd = struct.unpack(bo + "HI", raw[2:8])
    if magic != 42:
        raise ValueError("not a baseline TIFF")
    n, = struct.unpack(bo + "H", raw[ifd:ifd + 2])
    tags = {}
    for i in range(n):
        tag, typ, cnt, val = struct.unpack(bo + "HHI4s", raw[ifd + 2 + i * 12: ifd + 14 + i * 12])
        size = {1: 1, 2: 1, 3: 2, 4: 4}[typ] if typ in (1, 2, 3, 4) else 8
        code = {1: "B", 3: "H", 4: "I"}.get(typ)
        if code is None:
            continue
        if size * cnt <= 4:
            vals = struct.unpack(bo + "%d%s" % (cnt, code), val[:size * cnt])
        else:
            off, = struct.unpack(bo + "I", val)
            vals = struct.unpack(bo + "%d%s" % (cnt, code), raw[off: off + size * cnt])
        tags[tag] = vals
    w, h = tags[256][0], tags[257][0]
    c = tags.get(277, (1,))[0]
    if tags.get(259, (1,))[0] != 1 or tags.get(339, (1,))[0] != 3 or tags[258][0] != 32:
        raise ValueError("expected an uncompressed float32 TIFF")
    offs, cnts = tags[273], tags[279]
    data = b"".join(raw[o:o + k] for o, k in zip(offs, cnts))
    arr = np.frombuffer(data, dtype=bo + "f4", count=h * w * c).reshape(h, w, c)
    return arr.astype(np.float32)


def save_weights(dirpath: str, patch: np.ndarray, scale: float, mean_rgb, stddev_rgb) -> None:
    """attacker.py:328-341 (os.makedirs without exist_ok: an existing directory is an error there too)."""
    os.makedirs(dirpath)
    with open(os.path.join(dirpath, "scale.txt"), "w") as f:
        f.write(str(np.float32(scale)))
    img = np.clip(patch * np.float32(stddev_rgb) + np.float32(mean_rgb), 0.0, 255.0).astype(np.uint8)
    import cv2
    cv2.imwrite(os.path.join(dirpath, "patch.png"), img[..., ::-1])       # plt.imsave writes RGB; cv2 wants BGR
    write_tiff_f32(os.path.join(dirpath, "patch.tiff"), patch)


def load_weights(dirpath: str):
    """attacker.py:45-48."""
    patch = read_tiff_f32(os.path.join(dirpath, "patch.tiff"))
    with open(os.path.join(dirpath, "scale.txt")) as f:
        scale = ast.literal_eval(f.read())
    return patch, float(scale)
