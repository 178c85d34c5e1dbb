#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sorting.py tests/test_gpu_prescribed_motion.py tests/test_gpu_fsi.py -q -x -k "sort_system or reinit_density_matches or gate_kick or falling_spheres or young_modulus or falling_sphere_2d_kick or fsi_3d_kick or falling_water_column_fsi" > gpurun_out/r3z_sanitizer.log 2>&1; echo "rc=$?" >> gpurun_out/r3z_sanitizer.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|rc=" gpurun_out/r3z_sanitizer.log | tail -10
