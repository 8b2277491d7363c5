set -x
python -m pytest tests -m gpu -q > gpurun_out/r15_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r15_pytest_gpu.log; tail -3 gpurun_out/r15_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/r15_bench_ref.json 2> gpurun_out/r15_bench_ref.err
python bench.py > gpurun_out/r15_bench.json 2> gpurun_out/r15_bench.err; echo "bench rc=$?"
cat gpurun_out/r15_bench.json
