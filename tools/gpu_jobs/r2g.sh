#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/microbench/tmem_ld > gpurun_out/r2g_tmem_ld.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tmem_ld.log; cat gpurun_out/r2g_tmem_ld.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_bench_r2g.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interact_tiles -s 3 -c 1 -o gpurun_out/prof_r2g python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_full_r2g.log 2>&1; echo "ncu full rc=$?"
