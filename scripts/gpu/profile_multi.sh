# --set full capture of the 8-agent compact search kernel (75 776 trees x 2 000 iterations), after an un-profiled run
python scripts/sweep_multi.py 7 75776 2000 2 800 > gpurun_out/r2_multi_plain.log 2>&1; tail -1 gpurun_out/r2_multi_plain.log
ncu --set full --clock-control none --import-source on -k regex:search_compact_multi_kernel -c 1 -o gpurun_out/r2_search_compact_multi_v0 -f python scripts/sweep_multi.py 7 75776 2000 2 800 > gpurun_out/r2_ncu_multi.log 2>&1
echo done
