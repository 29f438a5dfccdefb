#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/exp10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/exp10_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/exp10_bench.json 2> gpurun_out/exp10_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/exp10_bench.err
python -c "
import json
r=json.loads(open('gpurun_out/exp10_bench.json').read().strip().splitlines()[-1])
print('value',round(r['value']),'frac',round(r['roofline']['frac'],4),'sec',round(r['roofline']['secondary']['frac'],3))
print('e2e',r['e2e'])
print('spot',r['parity_spot_check'])
print('workloads',{k:(round(v['value']),round(v['roofline_frac'],3)) for k,v in r['workloads'].items()})
print('shim',r.get('shim')); print('refgpu',r.get('reference_gpu_kernels')); print('reg',r.get('registration')); print('cpu',r['cpu_baseline'])
"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -5 | cut -c1-400
