timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_crowd_and_sizes.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
t() { echo -n "$*: "; env "$@" timeout 120 python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what fwd $SH 2>&1 | grep -E "us per call|rror" | tr '\n' ' '; echo; }
SH=""
for i in 1 2 3; do
t A=1
done
bash scripts/fwd_times.sh fwd 2>&1 | tail -6
