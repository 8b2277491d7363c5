python -m pytest tests/test_gpu_hypothesis_search.py -x -q 2>&1 | tail -2
echo "== diverged lanes"
python scripts/hyp_sweep.py "4096,10000;12288,10000" 2>&1 | grep -E "hyp|rror"
echo "== reconverged once per iteration + after the descent"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_hyprc.so python -m pytest tests/test_gpu_hypothesis_search.py -x -q 2>&1 | tail -2
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_hyprc.so python scripts/hyp_sweep.py "4096,10000;8192,10000;12288,10000" 2>&1 | grep -E "hyp|rror"
