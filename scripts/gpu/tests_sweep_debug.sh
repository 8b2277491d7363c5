python -m pytest tests -m gpu -q 2>&1 | tail -2
MAMCTS_SEARCH_DEBUG=1 python scripts/sweep_search.py "0,64,8192,1000,0,2,0;0,64,1024,10000,0,2,0;0,64,16384,10000,0,2,3200;0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw|blocks/SM" | sort -u
