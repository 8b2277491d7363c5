python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r106_bench_n2.json 2> gpurun_out/r106_bench_n2.err; echo "bench rc=$?"
head -c 300 gpurun_out/r106_bench_n2.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 scripts/check_multi_gpu.py 2>&1 | grep MULTI_GPU_CHECK
