python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --workload c3 2>&1 | grep -o '"env_steps_per_s": [0-9.]*'
python bench.py --workload c3 2>&1 | grep -o '"env_steps_per_s": [0-9.]*'
python scripts/sweep_search.py "32,64,65536,10000,0,2,3200" 2>&1 | grep lpw
