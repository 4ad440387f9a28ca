#!/bin/bash
# End-of-round profile of the apply kernels and of the bench step (each ncu pass only after its command exited 0 without ncu):
#   gpurun --timeout 500 -- 'bash scripts/profile_head.sh TAG'
# writes gpurun_out/prof_TAG.ncu-rep (--set full of one forward + backward call), gpurun_out/TAG_launches_bench.csv (launch list of
# the bench step, library kernels only) and gpurun_out/TAG_fwd_times.txt (per-kernel times of one call, caches as left).
tag=${1:-head}
mkdir -p gpurun_out
K='regex:^k_'
python scripts/kernel_loop.py --iters 1 --warmup 1 --what fwd,bwd > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 240 ncu --set full --clock-control none --import-source on -k "$K" -f -o gpurun_out/prof_${tag} \
    python scripts/kernel_loop.py --iters 1 --warmup 1 --what fwd,bwd > gpurun_out/${tag}_ncu_full.log 2>&1
echo "set full: $?"
bash scripts/fwd_times.sh fwd,bwd > gpurun_out/${tag}_fwd_times.txt 2>&1
echo "fwd_times: $?"
python bench.py --steps 2 --warmup 1 --launch-list --no-graphs > gpurun_out/${tag}_ll_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --launch-list --no-graphs > gpurun_out/${tag}_ncu_ll.log 2>&1
echo "launch list: $?"
