python -m pytest tests/test_gpu_hypothesis_search.py tests/test_gpu_search.py -x -q 2>&1 | tail -3
echo "== baseline (HEAD)"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_hypbase.so python scripts/hyp_sweep.py "4096,10000" 2>&1 | grep hyp
echo "== record prefetch"
python scripts/hyp_sweep.py "4096,10000;8192,10000;12288,10000" 2>&1 | grep -E "hyp|rror"
