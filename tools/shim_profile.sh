#!/bin/bash
# per-kernel GPU time of the drop-in class path:  gpurun -- 'bash tools/shim_profile.sh'
mkdir -p gpurun_out
python tools/shim_profile.py | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/shim_launches.csv pointcloud-depthfusion_b200/build/shim_fusion_test gpurun_out/shim_in.bin /tmp/shim_out.bin 10 > gpurun_out/shim_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/shim_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
body=rows[1:]
# the last 10 frames: take the tail third of the launches
from collections import OrderedDict
n=len(body); tail=body[-(n*10//31):] if n>62 else body
agg=OrderedDict()
for r in tail:
    k=r[ki][:48]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[vi])
tot=0
for k,(c,t) in agg.items(): print(f"{k:50s} n={c:3d} avg={t/c/1000:7.2f} us total={t/1000:8.1f} us"); tot+=t
print('launches in window', len(tail), 'GPU time', tot/1000, 'us')
PY
