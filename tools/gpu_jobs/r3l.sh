#!/bin/bash
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q -x > gpurun_out/r3l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3l_pytest_gpu.log; tail -6 gpurun_out/r3l_pytest_gpu.log
timeout 600 python tools/bench_fsi.py > gpurun_out/r3l_fsi_bench.txt 2> gpurun_out/r3l_fsi_bench.err; cat gpurun_out/r3l_fsi_bench.txt | cut -c1-420
for w in dam_break_2d dam_break_3d_small dam_break_3d_1m; do
    echo "== $w"; timeout 200 python bench.py --steps 30 --warmup 5 --quick --workload $w 2>&1 | tail -1 | cut -c1-330
done
