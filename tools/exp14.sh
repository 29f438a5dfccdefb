#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp14.jsonl; : > $O
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp14.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
for v in re0 re1 re1c3 re1s4; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
for sl in "6 3" "7 3" "8 4" "10 5"; do set -- $sl; run re1_s$1_l$2 DFB_B200_LIB=variants/re1/libdfb_b200.so DFB_STREAM_SLOTS=$1 DFB_STREAM_LAG=$2; done
DFB_B200_LIB=variants/re1i/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/exp14_instr.log 2>&1
grep -E "cta +(37|74) " gpurun_out/exp14_instr.log | sort | head -8 | cut -c1-300
DFB_B200_LIB=variants/re1/libdfb_b200.so timeout 600 python -m pytest tests/test_fuse_gpu.py -x -q -m gpu 2>&1 | tail -3
