python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --workload c3 --steps 20 --warmup 5 2>/dev/null | tail -c 1500
