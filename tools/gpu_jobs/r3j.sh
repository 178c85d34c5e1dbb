#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3j_subkey.txt; : > $out
for sk in 1 2 1 2; do
for sh in "" "--shuffle --sort"; do
  echo "== TPB_SUBKEY=$sk dam_break_3d_1m $sh" >> $out
  TPB_SUBKEY=$sk timeout 300 python bench.py --quick --workload dam_break_3d_1m --steps 30 --warmup 5 $sh 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(round(d.get('ms_per_step'),4), d.get('sort_ms'), {k: round(v,4) for k,v in d.get('phases_ms').items()})
" >> $out
done; done
cat $out
