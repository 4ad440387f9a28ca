timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|^eot' --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --launch-list --no-graphs > gpurun_out/ncu_bench.log 2>&1; echo ncu-bench rc=$?; tail -2 gpurun_out/ncu_bench.log
python scripts/launch_summary.py gpurun_out/r02_launches_bench.csv 2>&1 | tail -40
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_prepass|k_match|k_resize2|k_composite|k_bwd' -c 7 -o gpurun_out/prof_r02_s3_final -f python scripts/kernel_loop.py --iters 1 --warmup 0 --what fwd,bwd > gpurun_out/ncu_s3_final.log 2>&1; echo ncu-full rc=$?
timeout 300 python scripts/step_launch_list.py > gpurun_out/r02_step_launches.txt 2>&1; echo steplist rc=$?
bash scripts/fwd_times.sh fwd,bwd 2>&1 | tail -9
