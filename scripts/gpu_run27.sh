./oracle/_ref/dropin_test | tail -16
./oracle/_ref/dropin_bench 1000 2048 2>&1 | tee gpurun_out/r27_dropin_bench.log
