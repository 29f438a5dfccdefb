mkdir -p gpurun_out
python tools/reg_profile.py > gpurun_out/reg_profile.json 2> gpurun_out/reg_profile.err; cat gpurun_out/reg_profile.json; tail -3 gpurun_out/reg_profile.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/reg_launches.csv python tools/reg_profile.py --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/reg_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
from collections import OrderedDict
agg=OrderedDict()
for r in rows[1:]:
    k=r[ki][:40]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[vi])
for k,(n,t) in agg.items(): print(f"{k:42s} n={n:3d} total={t/1000:9.1f} us  avg={t/n/1000:7.2f} us")
PY
