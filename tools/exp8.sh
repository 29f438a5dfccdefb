#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp8.jsonl; : > $O
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp8.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
for v in v1 v1b c1b c2b; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
for v in c1 c2; do
for sl in "8 4" "10 5" "12 6" "10 3" "12 4"; do set -- $sl
run ${v}_s$1_l$2 DFB_B200_LIB=variants/$v/libdfb_b200.so DFB_STREAM_SLOTS=$1 DFB_STREAM_LAG=$2
done; done
run v1_s8 DFB_B200_LIB=variants/v1/libdfb_b200.so DFB_STREAM_SLOTS=8 DFB_STREAM_LAG=4
