#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r3v_large_single_gpu.txt; : > $out
free -g | head -2 >> $out
for wl in dam_break_3d_30m dam_break_3d_100m; do
  echo "== $wl" >> $out
  t0=$(date +%s)
  timeout 1200 python bench.py --quick --workload $wl --steps 5 --warmup 3 2> gpurun_out/r3v_$wl.err | tail -1 | cut -c1-700 >> $out
  echo "wall $(( $(date +%s) - t0 )) s" >> $out
  tail -2 gpurun_out/r3v_$wl.err >> $out
done
cat $out
