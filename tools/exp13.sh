#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp13.jsonl; : > $O
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/exp13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/exp13_pytest.log
Q="timeout 300 python tools/quick_bench.py --steps 6"
run() { tag=$1; shift; env "$@" $Q --tag $tag >> $O 2>> gpurun_out/exp13.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), r['env'])"; }
run final X=1
run final_s8 DFB_STREAM_SLOTS=8 DFB_STREAM_LAG=4
run exp_r2 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=2
run exp_r3 DFB_B200_LIB=variants/exp/libdfb_b200.so DFB_STREAM_REPEAT=3
DFB_B200_LIB=variants/instr/libdfb_b200.so DFB_STREAM_PROFILE=1 timeout 300 python tools/quick_bench.py --steps 1 --warmup 1 --framesets 128 --tag instr > gpurun_out/r02_instr.log 2>&1
grep -E "cta +(37|74) " gpurun_out/r02_instr.log | sort | head -8 | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --framesets 128 --no-cpu-baseline --no-e2e --no-workloads --no-shim"
$CMD > gpurun_out/r02_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fuse_stream -s 4 -c 1 -o gpurun_out/r02_stream $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -2
