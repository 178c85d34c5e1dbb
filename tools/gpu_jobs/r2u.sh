#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/run_fsi_hydrostatic_validation.py 0.35 > gpurun_out/r2u_fsi_validation.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_fsi_validation.log; tail -15 gpurun_out/r2u_fsi_validation.log
