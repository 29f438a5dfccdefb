#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp4.jsonl; : > $O
timeout 900 python -m pytest tests/test_fuse_gpu.py -x -q -m gpu > gpurun_out/exp4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/exp4_pytest.log
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp4.err; tail -1 $O | cut -c1-140; }
run v2 X=1
for v in w12 w20 s4 s8 nodeps; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
run exp_r2 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=2
run exp_r3 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=3
run lag2 DFB_STREAM_LAG=2
run lag4 DFB_STREAM_LAG=4
run slots7 DFB_STREAM_SLOTS=7 DFB_STREAM_LAG=3
DFB_B200_LIB=variants/instr/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/exp4_instr.log 2>&1
