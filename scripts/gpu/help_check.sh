python -m pytest tests/test_gpu_search.py tests/test_gpu_fuzz.py tests/test_gpu_dropin.py tests/test_gpu_pybind.py -x -q 2>&1 | tail -3
echo "== helpers on (default)"
python scripts/sweep_search.py "0,64,8192,1000,0,2,0;0,64,1024,10000,0,2,0;0,64,2368,10000,0,2,0;0,64,4096,1000,0,2,0" 2>&1 | grep -E "lpw"
echo "== helpers off"
MAMCTS_SEARCH_HELPERS=0 python scripts/sweep_search.py "0,64,8192,1000,0,2,0;0,64,1024,10000,0,2,0;0,64,2368,10000,0,2,0;0,64,4096,1000,0,2,0" 2>&1 | grep -E "lpw"
