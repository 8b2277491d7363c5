python -m pytest tests/test_gpu_search.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py tests/test_gpu_pybind.py tests/test_gpu_hypothesis_search.py -x -q 2>&1 | tail -3
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,75776,10000,0,2,3200;0,64,8192,1000,0,2,0" 2>&1 | grep -E "lpw"
