timeout 400 python -m pytest tests/test_gpu_forward.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
for i in 1 2; do for lib in mladversarialobjectdetection_b200/libeotpatch.so _ab/q2s2c4.so _ab/q2s3c3.so _ab/q1s2c7.so; do echo -n "$lib: "; EOTPATCH_LIB=$lib python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what fwd 2>&1 | grep "us per call"; done; done
