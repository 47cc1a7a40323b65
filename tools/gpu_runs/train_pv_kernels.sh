#!/bin/bash
cd /root/repo
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:"pv::" -c 3000 --csv \
    --log-file gpurun_out/launches_pv_train.csv python bench.py --workload train --steps 1 --warmup 0 > gpurun_out/bwd_ncu.log 2>&1
echo "exit $?"
python - <<'PY'
import csv, collections
rows=[l for l in open('gpurun_out/launches_pv_train.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for r in csv.DictReader(rows):
    if r['Metric Name']=='gpu__time_duration.sum':
        k=r['Kernel Name'].replace('void ','')[:58]
        v=float(r['Metric Value'].replace(',',''))
        if r['Metric Unit']=='us': v*=1000
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
print("total pv ms", round(tot/1e6,2))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print(f"{k:60s} n={a[0]:4d} total={a[1]/1e6:7.2f} ms avg={a[1]/a[0]/1e3:8.1f} us")
PY
