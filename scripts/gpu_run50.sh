python -m pytest tests/test_gpu_search.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for hot in 100000 1 16 64 256 1024; do
echo "== hot $hot"
MAMCTS_SEARCH_HOT=$hot python scripts/sweep_search.py "32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | grep lpw
done
