#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r4b_small_subkey.txt; : > $out
for rep in 1 2 3; do
for sk in 0 1; do
for wl in dam_break_2d dam_break_3d_small; do
  echo "== rep $rep TPB_SUBKEY=$sk $wl" >> $out
  TPB_SUBKEY=$sk timeout 200 python bench.py --quick --workload $wl --steps 200 --warmup 20 2>/dev/null | tail -1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(round(d.get('ms_per_step'),5), {k: round(v,4) for k,v in d.get('phases_ms').items()})
" >> $out
done; done; done
cat $out
