#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_gpu.log; tail -8 gpurun_out/r2h_pytest_gpu.log
for w in dam_break_3d_1m dam_break_3d_10m; do
  for e in 0 1; do
    if [ $e = 1 ]; then export TPB_SCAN3=1; else unset TPB_SCAN3; fi
    echo "== $w TPB_SCAN3=$e"; timeout 200 python bench.py --steps 20 --warmup 3 --quick --workload $w 2>&1 | tail -1 | cut -c1-420
  done
done
unset TPB_SCAN3
