for spw in 32 40 48 56; do
  echo "== async SPL=2 spw=$spw"
  MAMCTS_SEARCH_ASYNC=2 MAMCTS_SEARCH_ASYNC_SPW=$spw MAMCTS_SEARCH_DEBUG=1 timeout 300 python scripts/sweep_search.py "0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw|search_async" | sort | uniq
done
echo "== more trees, async SPL=2"
MAMCTS_SEARCH_ASYNC=2 MAMCTS_SEARCH_DEBUG=1 timeout 600 python scripts/sweep_search.py "0,64,113664,10000,0,2,3200;0,64,151552,10000,0,2,3200;0,64,151552,5000,0,2,1700" 2>&1 | grep -E "lpw|search_async|rror" | sort | uniq
echo "== more trees, lane per tree"
timeout 600 python scripts/sweep_search.py "0,64,151552,5000,0,2,1700;0,64,75776,5000,0,2,1700" 2>&1 | grep -E "lpw|rror"
