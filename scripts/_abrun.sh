timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_crowd_and_sizes.py tests/test_gpu_backward.py tests/test_gpu_fullsize.py tests/test_gpu_step.py -x -q 2>&1 | tail -3
t() { echo -n "$*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what fwd,bwd $SH 2>&1 | grep -E "us per call|rror" | tr '\n' ' '; echo; }
g() { echo -n "graph $*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --graph --iters 40 --warmup 3 --what fwd,bwd $SH 2>&1 | grep -E "us per call|rror" | tr '\n' ' '; echo; }
SH=""
t EOT_PDL=0
t EOT_PDL=1
t EOT_PDL=0
t EOT_PDL=1
g EOT_PDL=0
g EOT_PDL=1
SH="--batch 8 --image 1024 --patch 300"
t EOT_PDL=0
t EOT_PDL=1
g EOT_PDL=1
