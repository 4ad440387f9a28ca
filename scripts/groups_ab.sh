for g in 1 2 3 4 6 8; do echo "== G=$g"; EOT_FWD_GROUPS=$g python scripts/kernel_loop.py --time --graph --iters 30 --warmup 3 --what fwd 2>&1 | grep "us per call\|rror"; done
