"""A/B builds of libeotpatch.so with extra -D macros: `python scripts/ab_build.py NAME -DEOT_X=1 ...` -> _ab/NAME.so
(run a script against it with EOTPATCH_LIB=_ab/NAME.so)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mladversarialobjectdetection_b200.csrc import build as B

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "_ab")
obj_dir = os.path.join(out_dir, name + "_obj")
os.makedirs(obj_dir, exist_ok=True)
nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def one(item):
    src, flags = item
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    subprocess.run([nvcc, *B.NVCC_FLAGS, *flags, *extra, "-c", os.path.join(B.HERE, src), "-o", obj], check=True)
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(one, B.SOURCES.items()))
lib = os.path.join(out_dir, name + ".so")
subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs], check=True)
print(lib)
