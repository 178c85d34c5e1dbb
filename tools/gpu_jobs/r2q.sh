#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_gpu.log; tail -25 gpurun_out/r2t_pytest_gpu.log
