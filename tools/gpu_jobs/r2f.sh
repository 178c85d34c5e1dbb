#!/bin/bash
mkdir -p gpurun_out
timeout 60 tools/microbench/tc_filter > gpurun_out/r2f_tc_filter.log 2>&1; echo "tc rc=$?" >> gpurun_out/r2f_tc_filter.log; cat gpurun_out/r2f_tc_filter.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_gpu.log; tail -5 gpurun_out/r2f_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2f_bench.json; tail -3 gpurun_out/r2f_bench.err
