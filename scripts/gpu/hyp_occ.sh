echo "== 8 blocks/SM (default)"
python scripts/hyp_sweep.py "8192,10000;12288,10000" 2>&1 | grep -E "hyp|rror"
for mb in 10 12; do
  echo "== $mb blocks/SM"
  MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_hypmb$mb.so python scripts/hyp_sweep.py "8192,10000;12288,10000" 2>&1 | grep -E "hyp|rror"
done
echo "== 12 blocks/SM, lpw=2 at 8192 trees"
MAMCTS_SEARCH_LPW=2 MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_hypmb12.so python scripts/hyp_sweep.py "6144,10000;8192,10000" 2>&1 | grep -E "hyp|rror"
