ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/r23_rollout_c3_v2 -f python bench.py --workload c3 --steps 1 --warmup 1 > gpurun_out/r23_ncu_c3.log 2>&1
python bench.py > gpurun_out/r23_bench.json 2> gpurun_out/r23_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r23_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r23_launches.csv python bench.py --steps 2 --warmup 1 --cpu-repeats 1 > gpurun_out/r23_ncu_launches.log 2>&1
cat gpurun_out/r23_bench.json | head -c 3000
