set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r104_bench_n8.json 2> gpurun_out/r104_bench_n8.err; echo "bench rc=$?"
head -c 400 gpurun_out/r104_bench_n8.json; echo; tail -3 gpurun_out/r104_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 scripts/check_multi_gpu.py 2>&1 | grep MULTI_GPU_CHECK
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29539 bench.py --gpus 8 --trees 8192 --iterations 1000 --max-inner 0 --steps 5 --warmup 3 --cpu-repeats 1 > gpurun_out/r104_bench_c5_n8.json 2> gpurun_out/r104_bench_c5_n8.err; echo "c5 rc=$?"
head -c 400 gpurun_out/r104_bench_c5_n8.json; echo
