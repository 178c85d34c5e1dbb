#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prescribed_motion.py -q -x -k "gate" > gpurun_out/r3o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3o_pytest.log; tail -40 gpurun_out/r3o_pytest.log
