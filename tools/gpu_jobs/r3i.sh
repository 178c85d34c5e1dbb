#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sorting.py -q -x > gpurun_out/r3i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3i_pytest.log; tail -8 gpurun_out/r3i_pytest.log
out=gpurun_out/r3i_subkey.txt; : > $out
for sk in 0 1 2 3; do
for sh in "" "--shuffle" "--shuffle --sort" "--evolve 300"; do
  echo "== TPB_SUBKEY=$sk dam_break_3d_1m $sh" >> $out
  TPB_SUBKEY=$sk timeout 300 python bench.py --quick --workload dam_break_3d_1m --steps 30 --warmup 5 $sh 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(round(d.get('ms_per_step'),4), d.get('sort_ms'), {k: round(v,4) for k,v in d.get('phases_ms').items()})
" >> $out
done; done
for sk in 0 1; do
  echo "== TPB_SUBKEY=$sk dam_break_3d_10m" >> $out
  TPB_SUBKEY=$sk timeout 300 python bench.py --quick --workload dam_break_3d_10m --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(round(d.get('ms_per_step'),4), d.get('sort_ms'), {k: round(v,4) for k,v in d.get('phases_ms').items()})
" >> $out
done
cat $out
