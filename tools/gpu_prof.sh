#!/bin/bash
# one full ncu capture of a kernel: tools/gpu_prof.sh <tag> <kernel-regex> [extra bench args]
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --quick "$@" > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
