echo "== default build (8 blocks/SM), OVZ template + register diet"
MAMCTS_SEARCH_DEBUG=1 python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,8192,1000,0,2,0" 2>&1 | grep -E "lpw|blocks/SM" | sort | uniq
echo "== 10 blocks/SM (96 registers)"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_mb10.so MAMCTS_SEARCH_DEBUG=1 python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,94720,10000,0,2,3200" 2>&1 | grep -E "lpw|blocks/SM" | sort | uniq
