#!/bin/bash
# round 2 (b): full GPU test suite + default single-GPU bench line
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu) > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -5 gpurun_out/r2b_pytest_gpu.log
(time timeout 900 python bench.py --steps 50 --warmup 5) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 600 gpurun_out/r2b_bench.err
head -c 1500 gpurun_out/r2b_bench.json
