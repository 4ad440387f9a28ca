timeout 900 python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; echo bench rc=$?; tail -3 gpurun_out/bench_s3.err
