#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp7.jsonl; : > $O
timeout 600 python -m pytest tests/test_shim_gpu.py -x -q -m gpu > gpurun_out/exp7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/exp7_pytest.log
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp7.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
for v in v1n v1nf c1 c1s4 c1s5 c2 c3late c2late; do run $v DFB_B200_LIB=variants/$v/libdfb_b200.so; done
for v in v1n c1 c1s4 c2; do
run ${v}_slots7 DFB_B200_LIB=variants/$v/libdfb_b200.so DFB_STREAM_SLOTS=7 DFB_STREAM_LAG=3
run ${v}_slots8 DFB_B200_LIB=variants/$v/libdfb_b200.so DFB_STREAM_SLOTS=8 DFB_STREAM_LAG=4
done
run v1n_check DFB_B200_LIB=variants/v1n/libdfb_b200.so
timeout 300 python tools/quick_bench.py --steps 3 --check 4 --tag v1n_chk >> $O 2>> gpurun_out/exp7.err
pointcloud-depthfusion_b200/build/shim_fusion_test /tmp/in.bin /tmp/out.bin 2000 2>&1 | tail -2
