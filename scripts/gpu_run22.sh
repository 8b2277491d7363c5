for t in 1 2 4 6 8 12 16 24; do echo "service=$t"; MAMCTS_ROLLOUT_SERVICE=$t python bench.py --workload c3 2>&1 | grep -o '"env_steps_per_s": [0-9.]*'; done
