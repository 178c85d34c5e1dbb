#!/bin/bash
mkdir -p gpurun_out
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r3b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3b_pytest_gpu.log; tail -15 gpurun_out/r3b_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r3b_bench.json 2> gpurun_out/r3b_bench.err; tail -c 1500 gpurun_out/r3b_bench.json
