set -x
python -m pytest tests/test_gpu_search.py -m gpu -q -x > gpurun_out/r11_pytest_search.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r11_pytest_search.log
tail -15 gpurun_out/r11_pytest_search.log
python scripts/sweep_search.py "32,64,16384,10000,0,1,0;32,64,16384,10000,0,2,3200;32,64,32768,10000,0,2,3200;32,64,49152,10000,0,2,3200;32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | tee gpurun_out/r11_sweep.log
