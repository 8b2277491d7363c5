set -x
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
python scripts/sweep_search.py "32,64,16384,2000;16,64,16384,2000;8,64,16384,2000;4,64,16384,2000;8,32,16384,2000;16,32,16384,2000;32,64,32768,10000;16,64,32768,10000;8,64,32768,10000;8,64,21312,10000;16,64,42624,10000;4,64,32768,10000" > gpurun_out/r2_sweep.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu.log; tail -2 gpurun_out/r2_smoke.log; cat gpurun_out/r2_sweep.log
