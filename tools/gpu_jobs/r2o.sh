#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_end_slab.py 8 0.00271442 10 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -x -q 2>&1 | tail -3
for w in dam_break_3d_1m dam_break_3d_10m; do
    echo "== $w"; timeout 200 python bench.py --steps 20 --warmup 3 --quick --workload $w 2>&1 | tail -1 | cut -c1-330
done
