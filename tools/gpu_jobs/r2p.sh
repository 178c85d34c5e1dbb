#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_cells_tiles -s 2 -c 1 -o gpurun_out/prof_r2p_scan python tools/profile_end_slab.py 8 0.00271442 3 > gpurun_out/ncu_r2p.log 2>&1; echo "ncu rc=$?"
