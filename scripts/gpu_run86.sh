echo "== all child sectors prefetched"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_pfall.so python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,16384,10000,0,2,3200" 2>&1 | grep lpw
echo "== default"
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,16384,10000,0,2,3200" 2>&1 | grep lpw
