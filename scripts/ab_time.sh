#!/bin/bash
# A/B timing of eot_apply_fwd over several builds of the library, interleaved:  bash scripts/ab_time.sh ROUNDS lib.so ...
rounds=$1; shift
for r in $(seq $rounds); do
  for lib in "$@"; do
    echo -n "$lib: "
    EOTPATCH_LIB=$lib timeout 120 python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what ${WHAT:-fwd} 2>&1 | grep -E "us per call|rror" | tr '\n' ' '; echo
  done
done
