#!/bin/bash
# GPU experiment batch 1: where does the stream kernel's time go? (results -> gpurun_out/exp1.jsonl)
mkdir -p gpurun_out; O=gpurun_out/exp1.jsonl; : > $O
Q="python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp1.err; tail -1 $O | cut -c1-200; }
run base X=1
run base_check X=1 DFB_X=1
run formal DFB_B200_LIB=variants/formal/libdfb_b200.so
run exp_r1 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=1
run exp_r2 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=2
run exp_r3 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=3
run exp_r2_c2 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=2 DFB_STREAM_CTAS_PER_SM=2
run exp_r3_c2 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=3 DFB_STREAM_CTAS_PER_SM=2
run exp_r1_c1 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=1 DFB_STREAM_CTAS_PER_SM=1
run exp_r2_c1 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=2 DFB_STREAM_CTAS_PER_SM=1
run exp_r3_c1 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=3 DFB_STREAM_CTAS_PER_SM=1
run exp4_r1 DFB_B200_LIB=variants/exp4/libdfb_b200.so DFB_STREAM_REPEAT=1
run exp4_r2 DFB_B200_LIB=variants/exp4/libdfb_b200.so DFB_STREAM_REPEAT=2
run exp4_r3 DFB_B200_LIB=variants/exp4/libdfb_b200.so DFB_STREAM_REPEAT=3
# instrumented: cycle accounting of the three roles
DFB_B200_LIB=variants/instr/libdfb_b200.so DFB_STREAM_PROFILE=1 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/exp1_instr.log 2>&1
grep -c producer gpurun_out/exp1_instr.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> gpurun_out/exp1.err
