python -m pytest tests/test_gpu_hypothesis_search.py tests/test_gpu_dropin.py tests/test_gpu_pybind.py -x -q 2>&1 | tail -3
python scripts/hyp_sweep.py "4096,10000;8192,10000;12288,10000" 2>&1 | grep -E "hyp|rror"
