#!/bin/bash
cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"attn_bwd|kv_bwd_reduce|wgrad_partial" -c 60 --csv \
    --log-file gpurun_out/launches_bwd.csv python bench.py --workload train --steps 1 --warmup 0 > gpurun_out/bwd_ncu.log 2>&1
echo "exit $?"
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/launches_bwd.csv') if not l.startswith('==')]
import collections
agg=collections.OrderedDict()
for r in csv.DictReader(rows):
    if r['Metric Name']=='gpu__time_duration.sum':
        k=(r['Kernel Name'][:60], r['Grid Size'])
        v=float(r['Metric Value'].replace(',',''))
        if r['Metric Unit']=='us': v*=1000
        agg.setdefault(k,[]).append(v)
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1),'us')
PY
