for c in 0 25 44 58 72 100; do
echo "== preferred shared-memory carve-out $c %"
MAMCTS_SEARCH_CARVEOUT=$c MAMCTS_SEARCH_DEBUG=1 python scripts/sweep_search.py "0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw|blocks/SM" | sort -u
done
echo "== default"
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw"
