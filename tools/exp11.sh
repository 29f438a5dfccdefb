#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp11.jsonl; : > $O
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp11.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
for v in ra0 ra1 ra1_c2 ra1_c1 ra1_s3 ra1_w5 ra1_late; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
for sl in "6 3" "7 3" "8 3" "8 4" "10 5"; do set -- $sl; run ra1_s$1_l$2 DFB_B200_LIB=variants/ra1/libdfb_b200.so DFB_STREAM_SLOTS=$1 DFB_STREAM_LAG=$2; done
timeout 300 env DFB_B200_LIB=variants/ra1/libdfb_b200.so python tools/quick_bench.py --steps 3 --check 6 --tag ra1_chk | cut -c1-500
DFB_B200_LIB=variants/ra1/libdfb_b200.so timeout 600 python -m pytest tests/test_fuse_gpu.py -x -q -m gpu 2>&1 | tail -3
