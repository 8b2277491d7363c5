# async tree-pool scheduler: the whole GPU search test-suite with the pool kernel switched on, then A/B sweeps
for spl in 2 3 4; do
  echo "== tests with MAMCTS_SEARCH_ASYNC=$spl"
  MAMCTS_SEARCH_ASYNC=$spl timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4
done
echo "== baseline (lane per tree)"
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200" 2>&1 | grep lpw
for spl in 2 3 4; do
  echo "== async SPL=$spl"
  MAMCTS_SEARCH_ASYNC=$spl MAMCTS_SEARCH_DEBUG=1 timeout 300 python scripts/sweep_search.py "0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw|search_async" | sort | uniq
done
