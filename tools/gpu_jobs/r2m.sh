#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.log; tail -25 gpurun_out/r2m_pytest_gpu.log
for w in dam_break_3d_1m dam_break_3d_10m; do
    echo "== $w"; timeout 200 python bench.py --steps 20 --warmup 3 --quick --workload $w 2>&1 | tail -1 | cut -c1-420
done
timeout 300 python tools/profile_end_slab.py 8 0.00271442 10 2>&1 | tail -2
