ncu --set full --clock-control none --import-source on -k regex:search_compact_kernel -c 1 -o gpurun_out/r63_search_compact_v6 -f python bench.py --trees 16384 --iterations 2000 --steps 1 --warmup 1 --cpu-repeats 1 > gpurun_out/r63_ncu_search.log 2>&1
echo done
