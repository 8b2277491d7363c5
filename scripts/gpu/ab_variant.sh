V=${1:-pipe}
echo "== variant $V"
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_$V.so python -m pytest tests/test_gpu_search.py -m gpu -x -q 2>&1 | tail -1
MAMCTS_B200_LIB=$PWD/scripts/micro/libmamcts_$V.so python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,16384,10000,0,2,3200;0,64,1024,10000,0,2,0" 2>&1 | grep lpw
echo "== default"
python scripts/sweep_search.py "0,64,75776,10000,0,2,3200;0,64,16384,10000,0,2,3200;0,64,1024,10000,0,2,0" 2>&1 | grep lpw
