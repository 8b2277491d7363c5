ncu --set full --clock-control none --import-source on -k regex:search_compact_kernel -c 1 -o gpurun_out/r12_search_compact_v0 -f python bench.py --trees 16384 --iterations 2000 --layout 2 --steps 1 --warmup 1 --cpu-repeats 1 > gpurun_out/r12_ncu.log 2>&1
tail -3 gpurun_out/r12_ncu.log
