for i in 1 2; do for lib in mladversarialobjectdetection_b200/libeotpatch.so _ab/rm2.so _ab/rm8.so; do echo -n "$lib: "; EOTPATCH_LIB=$lib python scripts/kernel_loop.py --time --iters 40 --warmup 3 --what fwd 2>&1 | grep "us per call"; done; done
EOTPATCH_LIB=_ab/rm8.so bash scripts/fwd_times.sh fwd --warmup 3 | tail -3
