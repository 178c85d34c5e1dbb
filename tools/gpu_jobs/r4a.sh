#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/run_dam_break_validation.py > gpurun_out/r4a_dam_break_validation.log 2>&1; echo "rc=$?"; head -3 gpurun_out/r4a_dam_break_validation.log; tail -12 gpurun_out/r4a_dam_break_validation.log
