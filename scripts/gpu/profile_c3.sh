ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/r108_rollout_c3_v4 -f python bench.py --workload c3 --steps 1 --warmup 1 > gpurun_out/r108_ncu_c3.log 2>&1
echo done
