# --set full capture of search_compact_kernel at the C5 shape of one of 8 GPUs (8 192 trees x 1 000 iterations)
python scripts/sweep_search.py "0,64,8192,1000,0,2,0;0,64,32768,1000,0,2,0;0,64,65536,1000,0,2,0" > gpurun_out/r2_c5_plain.log 2>&1; cat gpurun_out/r2_c5_plain.log
ncu --set full --clock-control none --import-source on -k regex:search_compact_kernel -c 1 -o gpurun_out/r2_search_compact_c5shape -f python scripts/sweep_search.py "0,64,8192,1000,0,2,0" > gpurun_out/r2_ncu_c5.log 2>&1
echo done
