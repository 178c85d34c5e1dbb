#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sorting.py -q -x > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3h_pytest.log; tail -25 gpurun_out/r3h_pytest.log
out=gpurun_out/r3h_sort.txt; : > $out
for wl in dam_break_3d_1m dam_break_3d_10m; do
for sh in "" "--shuffle" "--shuffle --sort"; do
  echo "== $wl $sh" >> $out
  timeout 300 python bench.py --quick --workload $wl --steps 30 --warmup 5 $sh 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(d.get('ms_per_step'), d.get('sort_ms'), d.get('phases_ms'))
" >> $out
done; done
cat $out
