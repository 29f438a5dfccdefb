#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/exp15_bench2.json 2> gpurun_out/exp15_bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/exp15_bench2.err | cut -c1-300
python - <<'PY'
import json
try:
    r=json.loads([l for l in open('gpurun_out/exp15_bench2.json') if l.startswith('{')][-1])
    print('value',round(r['value']),'frac',round(r['roofline']['frac'],4))
    print('e2e',{k:v for k,v in r['e2e'].items() if k in ('value','link_gbs_per_gpu','limiter','copies_only')})
    print('gather',r.get('gather'))
    print('workloads',{k:(round(v['value']),round(v['roofline_frac'],3)) for k,v in r['workloads'].items()})
except Exception as e: print('ERR',e)
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) 2>&1 | grep -E "impl|real" | cut -c1-600
