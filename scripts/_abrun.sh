timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_crowd_and_sizes.py tests/test_gpu_fullsize.py tests/test_gpu_backward.py tests/test_gpu_variants.py -x -q 2>&1 | tail -3
EOTPATCH_LIB=_ab/geomdbg.so EOT_FWD_FUSED=1 EOT_FUSED_SKEW=32 EOT_KERNEL_TIMES=1 timeout 120 python scripts/kernel_loop.py --iters 1 --warmup 1 --what fwd 2>&1 | grep "slowest\|last geometry" | tail -2
t() { echo -n "$*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what fwd,bwd $SH 2>&1 | grep -E "us per call|rror" | tr '\n' ' '; echo; }
SH=""
for i in 1 2 3; do
t EOT_SORT_ITEMS=0
t EOT_SORT_ITEMS=1
done
bash scripts/fwd_times.sh fwd 2>&1 | tail -6
