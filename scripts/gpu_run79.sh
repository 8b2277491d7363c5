MAMCTS_SEARCH_DEBUG=1 MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_r112s4.so python scripts/sweep_search.py "0,64,85248,1000,0,2,2800" 2>&1 | grep -E "lpw|rror|blocks/SM" | head -3
MAMCTS_SEARCH_DEBUG=1 python scripts/sweep_search.py "0,64,75776,1000,0,2,2800" 2>&1 | grep -E "lpw|rror|blocks/SM" | head -3
