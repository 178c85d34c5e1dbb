#!/bin/bash
# A/B runs of library variants built into build_variants/lib_<name>.so:
#   tools/gpu_variants.sh "<names>" "<list lengths>"
mkdir -p gpurun_out
for v in $1; do
  cp build_variants/lib_$v.so trixiparticles.jl_b200/libtpb200.so
  echo "== variant $v"
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
  for L in $2; do
    echo "-- $v list $L"
    TPB_TILE_LIST_SPLIT=$L timeout 120 python bench.py --steps 20 --warmup 3 --quick 2>&1 | tail -1 | cut -c1-300
  done
done
