#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3r_xsplit.txt; : > $out
for xs in 1 2 3 4; do
  echo "== TPB_XSPLIT=$xs dam_break_3d_1m" >> $out
  TPB_XSPLIT=$xs timeout 300 python bench.py --quick --workload dam_break_3d_1m --steps 30 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(round(d.get('ms_per_step'),4), {k: round(v,4) for k,v in d.get('phases_ms').items()})
    else: print(line[:300])
" >> $out
done
cat $out
