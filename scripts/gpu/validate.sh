python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json;b=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]);print(b['value'],b['ms_per_step'],b['e2e']['value'],b['gpu_launches'],b['clocks'],b['leaf_batched']['env_steps_per_s'],b['rollout_c3']['env_steps_per_s'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 200; echo
