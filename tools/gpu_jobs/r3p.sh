#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fsi.py -q -x -k "falling_spheres or young_modulus_per" > gpurun_out/r3p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3p_pytest.log; tail -40 gpurun_out/r3p_pytest.log
