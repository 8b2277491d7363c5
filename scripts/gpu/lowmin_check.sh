python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,75776,10000,0,2,3200" 2>&1 | grep -E "lpw"
python bench.py --steps 5 --warmup 3 --cpu-repeats 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); c=d['rollout_c3']; print('c3', c['env_steps_per_s'], c['ms_per_decision'], c['variants']['long_episodes']['env_steps_per_s'], c['variants']['leaf_states']['env_steps_per_s']); print('c2', d['value'], d['ms_per_step'])"
