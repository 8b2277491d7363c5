set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r1_gpu.txt; nproc >> gpurun_out/r1_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1_smoke.log
python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench rc=$?" >> gpurun_out/r1_bench.err
python bench.py --impl reference > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
python bench.py --trees 16384 --iterations 2000 --steps 2 --warmup 1 --cpu-repeats 1 > gpurun_out/r1_bench_small.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --trees 16384 --iterations 2000 --steps 2 --warmup 1 --cpu-repeats 1 > gpurun_out/r1_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -c 1 -o gpurun_out/r1_search_v2 -f python bench.py --trees 16384 --iterations 2000 --steps 1 --warmup 1 --cpu-repeats 1 > gpurun_out/r1_ncu_full.log 2>&1
tail -3 gpurun_out/r1_pytest_gpu.log; cat gpurun_out/r1_smoke.log | tail -2; cat gpurun_out/r1_bench.json
