timeout 900 ncu --set full --clock-control none --import-source on -k regex:hyp_search_kernel -c 1 -o gpurun_out/${TAG:-a7}_hyp_search -f python scripts/hyp_sweep.py "4096,3000" > gpurun_out/${TAG:-a7}_ncu.log 2>&1
tail -3 gpurun_out/${TAG:-a7}_ncu.log
