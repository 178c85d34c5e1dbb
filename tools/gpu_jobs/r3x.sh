#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sorting.py -q -x -k "reinit" > gpurun_out/r3x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3x_pytest.log; tail -40 gpurun_out/r3x_pytest.log
