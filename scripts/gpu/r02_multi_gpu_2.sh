python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_ref_n2.json 2> gpurun_out/r02_ref_n2.err; echo "ref rc=$?"
tail -c 200 gpurun_out/r02_ref_n2.json
