ncu --set full --clock-control none --import-source on -k regex:search_compact_kernel -c 1 -o gpurun_out/r68_search_compact_v8 -f python bench.py --trees 75776 --iterations 2000 --steps 1 --warmup 1 --cpu-repeats 1 > gpurun_out/r68_ncu_search.log 2>&1
echo done
