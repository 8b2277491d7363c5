export MAMCTS_SEARCH_ASYNC=${SPL:-2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_async_kernel -c 1 -o gpurun_out/${TAG:-a3}_search_async -f python scripts/sweep_search.py "0,64,75776,2000,0,2,0" > gpurun_out/${TAG:-a3}_ncu.log 2>&1
tail -3 gpurun_out/${TAG:-a3}_ncu.log
