#!/bin/bash
# parity tests + a sweep of tile tuning parameters (quick device-resident timings)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in "114688 160" "114688 128" "114688 96" "75000 64" "75000 48" "56000 48" "56000 32"; do
  set -- $cfg
  TPB_TILE_SMEM=$1 TPB_TILE_LIST=$2 timeout 300 python bench.py --steps 20 --warmup 3 --quick 2>&1 | tail -1
done
python bench.py --steps 20 --warmup 3 --quick --variant 1 2>&1 | tail -1
