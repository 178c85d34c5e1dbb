#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "neighbor_sets or kick_dam_break_3d or chunked or stress" 2>&1 | tail -2
for w in dam_break_3d_1m dam_break_3d_10m; do
    echo "== $w"; timeout 200 python bench.py --steps 20 --warmup 3 --quick --workload $w 2>&1 | tail -1 | cut -c1-330
done
