#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fsi.py -q -x -s -k "oscillating_beam" > gpurun_out/r3n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3n_pytest.log; tail -30 gpurun_out/r3n_pytest.log
