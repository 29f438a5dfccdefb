#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp21.jsonl; : > $O
timeout 900 python -m pytest tests/test_fuse_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { tag=$1; shift; timeout 300 python tools/quick_bench.py --steps 6 --tag $tag "$@" >> $O 2>> gpurun_out/exp21.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
run final; run final_again
DFB_B200_LIB=variants/s4/libdfb_b200.so run s4
DFB_STREAM_SLOTS=6 DFB_STREAM_LAG=3 run slots6
DFB_STREAM_SLOTS=8 DFB_STREAM_LAG=4 run slots8
run C1 --workload C1; run C4 --workload C4; run C5 --workload C5
DFB_B200_LIB=variants/instr/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/exp21_instr.log 2>&1
grep -E "cta +(37|74) " gpurun_out/exp21_instr.log | sort | head -8 | cut -c1-300
