python -m pytest tests/test_gpu_rollout.py -m gpu -q -x 2>&1 | tail -3
python bench.py --workload c3 2>&1 | head -c 420; echo
python bench.py --workload c3 2>&1 | head -c 420; echo
