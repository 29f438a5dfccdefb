#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp9.jsonl; : > $O
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp9.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
for v in v1b v1bf v1like v1like_ahead v1like_s3 c2st c2st_ahead c1st c1b c2b; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
run v1like_chk DFB_B200_LIB=variants/v1like/libdfb_b200.so
timeout 300 env DFB_B200_LIB=variants/v1like/libdfb_b200.so python tools/quick_bench.py --steps 3 --check 4 --tag v1like_chk | cut -c1-400
run v1like_dyn DFB_B200_LIB=variants/v1like_dyn/libdfb_b200.so
run v1like_ahead_s8 DFB_B200_LIB=variants/v1like_ahead/libdfb_b200.so DFB_STREAM_SLOTS=8 DFB_STREAM_LAG=4
Q="timeout 300 python tools/quick_bench.py --steps 4 --workload C5"
run C5_default DFB_B200_LIB=variants/v1like/libdfb_b200.so
run C5_s1 DFB_B200_LIB=variants/v1like/libdfb_b200.so DFB_STREAM_SLOTS=1 DFB_STREAM_LAG=0
run C5_c1b_s1 DFB_B200_LIB=variants/c1b/libdfb_b200.so DFB_STREAM_SLOTS=1 DFB_STREAM_LAG=0
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_shim_gpu.py tests/test_fuse_gpu.py -x -q -m gpu 2>&1 | tail -5
