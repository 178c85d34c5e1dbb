#!/bin/bash
mkdir -p gpurun_out

timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2l_endslab.csv python tools/profile_end_slab.py 8 0.00271442 3 > gpurun_out/ncu_r2l.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_r2l_endslab.csv') if l.startswith('"')))
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); iid=h.index("ID")
d=collections.defaultdict(dict)
for r in rows[1:]:
    d[(r[iid],r[ik].split('(')[0][:60])][r[im]]=float(r[iv].replace(',',''))
agg=collections.defaultdict(list)
for (i,k),m in d.items(): agg[k].append(m)
for k,v in agg.items():
    v=v[len(v)//2:]
    t=sorted(x['gpu__time_duration.sum'] for x in v)[len(v)//2]
    rd=sorted(x.get('dram__bytes_read.sum',0) for x in v)[len(v)//2]; wr=sorted(x.get('dram__bytes_write.sum',0) for x in v)[len(v)//2]
    print(f"{k:62s} n={len(v):3d} t={t/1e3:9.1f} us  dram rd {rd/1e6:8.1f} MB wr {wr/1e6:8.1f} MB")
PY
