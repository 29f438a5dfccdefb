#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp3.jsonl; : > $O
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp3.err; tail -1 $O | cut -c1-140; }
for v in nofence batch4 batch8 late late_s3 s2 nodeps nodeps_nofence; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
run v2_slots5 DFB_STREAM_SLOTS=5 DFB_STREAM_LAG=2
run v2_slots7 DFB_STREAM_SLOTS=7 DFB_STREAM_LAG=3
for v in instr nofence_i; do
DFB_B200_LIB=variants/$v/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag $v > gpurun_out/exp3_$v.log 2>&1
done
