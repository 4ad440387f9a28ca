timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_crowd_and_sizes.py tests/test_gpu_step.py -x -q 2>&1 | tail -3
t() { echo -n "$*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what bwd 2>&1 | grep -E "us per call|rror" | head -3; }
for i in 1 2; do
t EOTPATCH_LIB=_ab/base.so
t A=1
done
EOT_KERNEL_TIMES=1 timeout 120 python scripts/kernel_loop.py --iters 3 --warmup 1 --what bwd 2>&1 | grep eot | tail -2
EOTPATCH_LIB=_ab/base.so EOT_KERNEL_TIMES=1 timeout 120 python scripts/kernel_loop.py --iters 3 --warmup 1 --what bwd 2>&1 | grep eot | tail -2
