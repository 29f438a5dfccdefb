#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/exp16_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/exp16_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/exp16_bench.json 2> gpurun_out/exp16_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/exp16_bench.err
python -c "
import json
r=json.loads(open('gpurun_out/exp16_bench.json').read().strip().splitlines()[-1])
print('value',round(r['value']),'frac',round(r['roofline']['frac'],4),'sec',round(r['roofline']['secondary']['frac'],3))
print('e2e',r['e2e']['value'], r['e2e']['copies_only']['value'], r['e2e']['limiter'])
print('spot',r['parity_spot_check']['mismatches'])
print('workloads',{k:(round(v['value']),round(v['roofline_frac'],3)) for k,v in r['workloads'].items()})
print('shim',r.get('shim',{}).get('value'), r.get('shim',{}).get('vs_reference_gpu_kernels')); print('refgpu',r.get('reference_gpu_kernels',{}).get('value'));
"
CMD="python bench.py --steps 2 --warmup 3 --framesets 128 --no-cpu-baseline --no-e2e --no-workloads --no-shim"
$CMD > gpurun_out/r02_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fuse_stream -s 4 -c 1 -o gpurun_out/r02_stream_b $CMD > gpurun_out/r02_ncu_full.log 2>&1
echo "full capture rc=$?"
