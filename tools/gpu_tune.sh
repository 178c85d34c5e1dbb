#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 3 --quick 2>&1 | tail -1
TPB_TILE_SMEM=75000 TPB_TILE_LIST=64 timeout 300 python bench.py --steps 20 --warmup 3 --quick 2>&1 | tail -1
bash tools/gpu_prof.sh tiles3 k_interact_tiles
