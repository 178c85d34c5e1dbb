#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/run_validation_resolutions.py > gpurun_out/r3s_validation.txt 2> gpurun_out/r3s_validation.err; echo "rc=$?"; cat gpurun_out/r3s_validation.txt; tail -5 gpurun_out/r3s_validation.err
