#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp18.jsonl; : > $O
timeout 900 python -m pytest tests/test_fuse_gpu.py -x -q -m gpu 2>&1 | tail -4
run() { tag=$1; shift; timeout 300 python tools/quick_bench.py --steps 5 --tag $tag "$@" >> $O 2>> gpurun_out/exp18.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), 'covered', round(r['covered_fraction'],4), r['env'])"; }
run base
for k in 1 4 8 16; do DFB_STREAM_AGGREGATE=1 run agg_shrink$k --out-shrink $k; DFB_STREAM_AGGREGATE=0 run noagg_shrink$k --out-shrink $k; done
run auto_shrink8 --out-shrink 8
run auto_shrink4 --out-shrink 4
