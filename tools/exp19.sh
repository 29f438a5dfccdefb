#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/exp19_bench$N.json 2> gpurun_out/exp19_bench$N.err; echo "bench$N rc=$?"; tail -5 gpurun_out/exp19_bench$N.err | cut -c1-300
python - <<PY
import json
try:
    r=json.loads([l for l in open('gpurun_out/exp19_bench$N.json') if l.startswith('{')][-1])
    print('value',round(r['value']),'frac',round(r['roofline']['frac'],4))
    print('e2e',{k:v for k,v in r['e2e'].items() if k in ('value','link_gbs_per_gpu','limiter')}, r['e2e']['copies_only']['value'])
    g=r.get('gather'); print('gather',{k:g[k] for k in ('value','rank0_ingress_gbs','matches_single_gpu','nccl_gather_value')})
    print('workloads',{k:(round(v['value']),round(v['roofline_frac'],3)) for k,v in r['workloads'].items()})
except Exception as e: print('ERR',e)
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) 2>&1 | grep -E "impl|real" | cut -c1-250
nvidia-smi topo -m > gpurun_out/r02_topo$N.txt 2>&1
