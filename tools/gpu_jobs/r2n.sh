#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_end_slab.py 8 0.00271442 10 2>&1 | tail -1
TPB_SCAN3=1 timeout 300 python tools/profile_end_slab.py 8 0.00271442 10 2>&1 | tail -1
TPB_SCAN3=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches_r2n_scan3.csv python tools/profile_end_slab.py 8 0.00271442 3 > gpurun_out/ncu_r2n.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_r2n_scan3.csv') if l.startswith('"')))
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[ik].split('(')[0][:60]].append(float(r[iv].replace(',','')))
for k,v in agg.items():
    v=v[len(v)//2:]; print(f"{k:62s} n={len(v):3d} t={sorted(v)[len(v)//2]/1e3:9.1f} us")
PY
