python -m pytest tests/test_gpu_search.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
python scripts/sweep_search.py "32,64,75776,10000,0,2,3200;32,64,65536,10000,0,2,3200" 2>&1 | grep lpw
