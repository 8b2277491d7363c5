echo "== C5 shape per GPU: 8192 trees x 1000 iterations; lanes per warp sweep (block 64 and 32)"
python scripts/sweep_search.py "32,64,8192,1000,0,2,0;16,64,8192,1000,0,2,0;8,64,8192,1000,0,2,0;4,64,8192,1000,0,2,0;2,64,8192,1000,0,2,0;1,64,8192,1000,0,2,0;8,32,8192,1000,0,2,0;4,32,8192,1000,0,2,0" 2>&1 | grep lpw
echo "== 1024 trees x 10000 iterations"
python scripts/sweep_search.py "32,64,1024,10000,0,2,0;8,64,1024,10000,0,2,0;2,64,1024,10000,0,2,0;1,64,1024,10000,0,2,0;1,32,1024,10000,0,2,0" 2>&1 | grep lpw
echo "== 16384 trees x 10000 iterations"
python scripts/sweep_search.py "32,64,16384,10000,0,2,3200;16,64,16384,10000,0,2,3200;8,64,16384,10000,0,2,3200;4,64,16384,10000,0,2,3200" 2>&1 | grep lpw
echo "== 32768 trees x 10000 iterations"
python scripts/sweep_search.py "32,64,32768,10000,0,2,3200;16,64,32768,10000,0,2,3200;8,64,32768,10000,0,2,3200" 2>&1 | grep lpw
