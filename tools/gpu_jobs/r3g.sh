#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3g_shuffle.txt; : > $out
for wl in dam_break_3d_1m dam_break_3d_10m; do
for sh in "" "--shuffle"; do
  echo "== $wl $sh" >> $out
  timeout 300 python bench.py --quick --workload $wl --steps 30 --warmup 5 $sh 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d.get('ms_per_step'), d.get('phases_ms'))
" >> $out
done; done
cat $out
