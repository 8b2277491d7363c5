python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/sweep_search.py "32,64,16384,10000,0,2,3200;32,64,32768,10000,0,2,3200;32,64,49152,10000,0,2,3200;32,64,65536,10000,0,2,3200;32,64,75776,10000,0,2,3200" 2>&1 | tee gpurun_out/r17_sweep.log
