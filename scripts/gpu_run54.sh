for v in 6 8 10; do
echo "== path levels in shared memory: $v"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_sh$v.so python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -1
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_sh$v.so python scripts/sweep_search.py "32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | grep lpw
done
