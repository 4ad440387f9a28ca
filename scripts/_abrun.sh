nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_2gpu.json 2> gpurun_out/bench_c2_2gpu.err; echo c2x2 rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; echo c3x2 rc=$?
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_1gpu.json 2> gpurun_out/bench_c4_1gpu.err; echo c4 rc=$?
