#!/bin/bash
# stage times of eot_apply_fwd (CUDA events between the kernels, warm caches): bash scripts/stage_times.sh [lib.so ...]
for lib in "${@:-mladversarialobjectdetection_b200/libeotpatch.so}"; do
  echo "== $lib"
  EOT_KERNEL_TIMES=1 EOTPATCH_LIB=$lib python scripts/kernel_loop.py --iters 6 --warmup 2 --what ${WHAT:-fwd} 2>&1 | grep "^\[eot\]" | tail -4
  EOTPATCH_LIB=$lib python scripts/kernel_loop.py --time --iters 30 --warmup 3 --what ${WHAT:-fwd} 2>&1 | grep "us per call"
done
