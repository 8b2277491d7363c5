timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_search.py -m gpu -q -x -k "not large_trees and not statistically" > gpurun_out/r25_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/r25_memcheck.log
