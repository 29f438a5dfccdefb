#!/bin/bash
# Ring-parameter sweep of the stream kernel (DFB_STREAM_SLOTS / DFB_STREAM_LAG) per workload, on the GPU box:
#   gpurun -- 'bash tools/ring_sweep.sh'      -> gpurun_out/ring_sweep.jsonl (one quick_bench line per point)
mkdir -p gpurun_out; O=gpurun_out/ring_sweep.jsonl; : > $O
run() { wl=$1; s=$2; l=$3; shift 3
  DFB_STREAM_SLOTS=$s DFB_STREAM_LAG=$l timeout 300 python tools/quick_bench.py --workload $wl --steps 6 --tag ${wl}_s${s}_l${l} "$@" >> $O 2>> gpurun_out/ring_sweep.err
  tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), round(r['frac'],4))"; }
for sl in "6 3" "7 3" "7 2" "7 4" "8 4" "8 3"; do run C2 $sl; done
for sl in "5 2" "6 3" "7 3" "7 2" "8 4" "8 3" "9 4"; do run C4 $sl; done
for sl in "2 1" "3 1" "3 2" "4 2" "1 0"; do run C5 $sl; done
for sl in "12 6" "16 8" "16 6" "16 10" "20 10" "24 12"; do run C1 $sl; done
