#!/bin/bash
# round 2 (c): FSI tests
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_fsi.py -x -q -m gpu) > gpurun_out/r2c_fsi.log 2>&1
tail -30 gpurun_out/r2c_fsi.log
