#!/bin/bash
# GPU experiment batch 2: first run of the fat-CTA stream kernel
mkdir -p gpurun_out; O=gpurun_out/exp2.jsonl; : > $O
timeout 900 python -m pytest tests/test_fuse_gpu.py -x -q -m gpu > gpurun_out/exp2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/exp2_pytest.log
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp2.err; tail -1 $O | cut -c1-160; }
run v2 X=1
run v2_check X=1
run v1 DFB_B200_LIB=variants/v1/libdfb_b200.so
run w12 DFB_B200_LIB=variants/w12/libdfb_b200.so
run w20 DFB_B200_LIB=variants/w20/libdfb_b200.so
run w16s4 DFB_B200_LIB=variants/w16s4/libdfb_b200.so
run w16s8 DFB_B200_LIB=variants/w16s8/libdfb_b200.so
run exp_r2 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=2
run exp_r3 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=3
run v2_slots8 DFB_STREAM_SLOTS=8 DFB_STREAM_LAG=4
run v2_lag2 DFB_STREAM_LAG=2
run v2_lag4 DFB_STREAM_LAG=4
DFB_B200_LIB=variants/instr/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/exp2_instr.log 2>&1
timeout 600 python tools/quick_bench.py --steps 3 --check 6 --tag check512 >> $O 2>> gpurun_out/exp2.err; tail -1 $O | cut -c1-400
