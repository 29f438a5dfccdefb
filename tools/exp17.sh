#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out/exp17.jsonl; : > $O
run() { tag=$1; shift; timeout 300 python tools/quick_bench.py --steps 5 --tag $tag "$@" >> $O 2>> gpurun_out/exp17.err; tail -1 $O | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['tag'], round(r['framesets_per_s']), 'covered', round(r['covered_fraction'],4))"; }
for k in 1 2 4 8 16; do run shrink$k --out-shrink $k; done
for k in 1 4 16; do run h_shrink$k --out-shrink $k --horizontal; done
run sparse --density sparse
