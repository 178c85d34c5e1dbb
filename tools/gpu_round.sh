#!/bin/bash
# GPU round: parity tests, bench, ncu launch list, one full capture of the top kernel.
# usage: tools/gpu_round.sh [tag] [kernel-regex]
TAG=${1:-r}
KREGEX=${2:-k_interact_tiles}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
