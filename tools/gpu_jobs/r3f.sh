#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_fsi.py > gpurun_out/r3f_fsi_bench.txt 2> gpurun_out/r3f_fsi_bench.err; echo "rc=$?"; cat gpurun_out/r3f_fsi_bench.txt; tail -5 gpurun_out/r3f_fsi_bench.err
