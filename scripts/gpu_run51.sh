python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -2
echo "== default lib (policy operand, all normal)"
python scripts/sweep_search.py "32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | grep lpw
for v in 0 1 2; do
echo "== evict_first from depth $v"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_cold$v.so python scripts/sweep_search.py "32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | grep lpw
done
