python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --workload c3 2>&1 | head -c 400
