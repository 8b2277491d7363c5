( time ./oracle/_ref/dropin_hypothesis ) 2>&1 | tail -20
python -m pytest tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4
