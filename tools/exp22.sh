#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp22.jsonl; : > $O
run() { tag=$1; shift; timeout 300 python tools/quick_bench.py --steps 6 --tag $tag "$@" >> $O 2>> gpurun_out/exp22.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
run final
DFB_B200_LIB=variants/c3w8s2/libdfb_b200.so run c3w8s2
DFB_B200_LIB=variants/c3w8s2/libdfb_b200.so DFB_STREAM_SLOTS=9 DFB_STREAM_LAG=4 run c3w8s2_s9
DFB_B200_LIB=variants/c3w8s2/libdfb_b200.so run c3w8s2_h --horizontal
run final_h --horizontal
# memcheck on a small ring-dependency case (one tool, small workload)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_fuse_gpu.py::test_stream_kernel_ring_dependencies[3-1]" "tests/test_fuse_gpu.py::test_fused_matches_oracle[3cam-small-roi]" -x -q -m gpu > gpurun_out/exp22_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/exp22_memcheck.log | cut -c1-200
