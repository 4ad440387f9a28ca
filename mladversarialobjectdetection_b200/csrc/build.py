"""Builds libeotpatch.so in-tree with nvcc for sm_100a (no torch extension machinery: the product
boundary is a plain C ABI).  -fmad=false: see the parity note in eot_common.cuh."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libeotpatch.so")
SOURCES = ["capi.cu", "eot_fwd.cu", "eot_bwd.cu", "score_max.cu", "patch_opt.cu"]
HEADERS = ["eot_common.cuh", os.path.join("..", "..", "include", "eotpatch.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "--expt-relaxed-constexpr"]


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        p = os.path.join(HERE, f)
        if os.path.exists(p):
            h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = LIB + ".stamp"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(HERE, f) for f in SOURCES if os.path.exists(os.path.join(HERE, f))]
    cmd = [nvcc, *NVCC_FLAGS, "-shared", "-o", LIB, *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
