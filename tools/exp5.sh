#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp5.jsonl; : > $O
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_shim_gpu.py -x -q -m gpu > gpurun_out/exp5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/exp5_pytest.log
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp5.err; tail -1 $O | cut -c1-140; }
run default X=1
for v in c1 c2 c3 c2s3 c2w10 c3s3; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
DFB_B200_LIB=variants/c2i/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/exp5_instr.log 2>&1
run c2_slots7 DFB_B200_LIB=variants/c2/libdfb_b200.so DFB_STREAM_SLOTS=7 DFB_STREAM_LAG=3
run c2_C1 DFB_B200_LIB=variants/c2/libdfb_b200.so
