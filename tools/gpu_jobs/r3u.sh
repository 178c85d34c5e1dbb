#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_time_loop.py -q -x -k "hydrostatic_water_column_3d" > gpurun_out/r3u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3u_pytest.log; tail -30 gpurun_out/r3u_pytest.log
