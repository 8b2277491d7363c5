python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,8192,1000,0,2,0" 2>&1 | grep -E "lpw"
python bench.py --workload c3 --steps 5 --warmup 3 --cpu-repeats 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('c3', d.get('value'), d.get('ms_per_step'))"
