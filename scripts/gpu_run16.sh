set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/check_multi_gpu.py > gpurun_out/r16_multi_check.log 2>&1; echo "check rc=$?"
grep MULTI_GPU_CHECK gpurun_out/r16_multi_check.log || tail -20 gpurun_out/r16_multi_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r16_bench_n2.json 2> gpurun_out/r16_bench_n2.err; echo "bench rc=$?"
cat gpurun_out/r16_bench_n2.json | head -c 1500; tail -5 gpurun_out/r16_bench_n2.err
