#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3d_adami_grid.txt; : > $out
for wl in dam_break_3d_1m dam_break_3d_10m; do
for g in 0 296 444 592 740 888 1184 2368; do
  echo "== $wl TPB_ADAMI_GRID=$g" >> $out
  TPB_ADAMI_GRID=$g timeout 300 python bench.py --quick --workload $wl --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d.get('ms_per_step'), d.get('phases_ms') or d.get('config'))
" >> $out
done; done
cat $out
