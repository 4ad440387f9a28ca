t() { echo -n "$*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what fwd 2>&1 | grep -E "us per call|rror" | head -3; }
g() { echo -n "graph $*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --graph --iters 40 --warmup 3 --what fwd 2>&1 | grep -E "us per call|rror" | head -3; }
t EOT_FWD_CE=0
t EOT_FWD_CE=2
t EOT_FWD_CE=2 EOT_PREPASS_BULK_ON=0
g EOT_FWD_CE=2
EOT_FWD_CE=2 EOT_KERNEL_TIMES=1 timeout 120 python scripts/kernel_loop.py --iters 3 --warmup 1 --what fwd 2>&1 | grep eot | tail -2
EOT_FWD_CE=2 EOT_PREPASS_BULK_ON=0 EOT_KERNEL_TIMES=1 timeout 120 python scripts/kernel_loop.py --iters 3 --warmup 1 --what fwd 2>&1 | grep eot | tail -2
