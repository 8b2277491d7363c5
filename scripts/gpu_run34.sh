python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/sweep_search.py "32,64,32768,10000,0,1,0" 2>&1 | grep lpw
