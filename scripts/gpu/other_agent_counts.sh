python -m pytest tests -m gpu -q 2>&1 | tail -2
echo "== 8 agents (classic layout, hash)"
SWEEP_OTHERS=7 python scripts/sweep_search.py "0,64,8192,10000;0,64,16384,10000" 2>&1 | grep lpw
echo "== 3 agents"
SWEEP_OTHERS=2 python scripts/sweep_search.py "0,64,16384,10000;0,64,32768,10000" 2>&1 | grep lpw
echo "== 2 agents classic"
python scripts/sweep_search.py "0,64,32768,10000,0,1,0" 2>&1 | grep lpw
