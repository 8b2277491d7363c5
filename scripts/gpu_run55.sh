for v in 2 3 4 5; do
echo "== path levels in shared memory: $v"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_sh$v.so python scripts/sweep_search.py "32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | grep lpw
done
