echo "== default"
python scripts/sweep_search.py "32,64,75776,10000,0,2,3200" 2>&1 | grep lpw
echo "== all sectors of the child table prefetched with the record"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_pfall.so python scripts/sweep_search.py "32,64,75776,10000,0,2,3200;32,64,65536,10000,0,2,3200" 2>&1 | grep lpw
echo "== 10 blocks per SM (96 registers, spills)"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_mb10.so python scripts/sweep_search.py "32,64,75776,10000,0,2,3200;32,64,94720,10000,0,2,3200" 2>&1 | grep lpw
