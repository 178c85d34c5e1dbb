#!/bin/bash
# round 2 (d): new tests (integrators, parity holes)
mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests/test_gpu_time_loop.py tests/test_gpu_parity.py -x -q -m gpu -k "new_device_vector or rdpk3 or shipped_integrator or bench_workload or 3d_bruteforce or summation_density_3d") > gpurun_out/r2d_tests.log 2>&1
tail -40 gpurun_out/r2d_tests.log
