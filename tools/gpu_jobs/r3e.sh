#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_prescribed_motion.py tests/test_gpu_fsi.py -q -x -k "accelerated_tank_kick or moving_wall_kick or moving_clamped or split_integration_pieces" > gpurun_out/r3e_sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/r3e_sanitizer.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|rc=" gpurun_out/r3e_sanitizer.log | tail -10
