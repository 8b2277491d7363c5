python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r58_launches.csv python bench.py --steps 2 --warmup 1 --cpu-repeats 1 > gpurun_out/r58_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_compact_kernel -c 1 -o gpurun_out/r58_search_compact_v5 -f python bench.py --trees 16384 --iterations 2000 --steps 1 --warmup 1 --cpu-repeats 1 > gpurun_out/r58_ncu_search.log 2>&1
echo done
