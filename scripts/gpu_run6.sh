set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r6_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r6_pytest_gpu.log
tail -3 gpurun_out/r6_pytest_gpu.log
python scripts/sweep_search.py "32,64,16384,2000;32,64,32768,10000;32,64,49152,10000" 2>&1 | tee gpurun_out/r6_sweep.log
python bench.py --workload c3 > gpurun_out/r6_c3.json 2>&1; cat gpurun_out/r6_c3.json | head -c 500
ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o gpurun_out/r6_search_v4 -f python bench.py --trees 16384 --iterations 2000 --steps 1 --warmup 1 --cpu-repeats 1 > gpurun_out/r6_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/r6_rollout_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 > gpurun_out/r6_ncu_c3.log 2>&1
