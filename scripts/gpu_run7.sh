for g in 32 64 128; do
  echo "== L2 fetch granularity $g"
  MAMCTS_L2_FETCH_GRANULARITY=$g python scripts/sweep_search.py "32,64,32768,10000;32,64,49152,10000"
done 2>&1 | tee gpurun_out/r7_l2gran.log
