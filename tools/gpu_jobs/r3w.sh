#!/bin/bash
mkdir -p gpurun_out
for sk in 26 27; do
  TPB_SUBKEY=$sk timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sorting.py tests/test_gpu_slabs.py -m gpu -q -x -k "kick_dam_break_3d or neighbor_sets or sort_system or slab_ranks or chunked or random_clustered" > gpurun_out/r3w_pytest_$sk.log 2>&1; echo "TPB_SUBKEY=$sk pytest rc=$?"; tail -2 gpurun_out/r3w_pytest_$sk.log
done
out=gpurun_out/r3w_keybits.txt; : > $out
for sk in 0 1 26 27; do
  echo "== TPB_SUBKEY=$sk dam_break_3d_1m --shuffle" >> $out
  TPB_SUBKEY=$sk timeout 300 python bench.py --quick --workload dam_break_3d_1m --steps 30 --warmup 5 --shuffle 2>/dev/null | tail -1 | cut -c1-330 >> $out
done
for sk in 0 1; do
  echo "== TPB_SUBKEY=$sk dam_break_3d_100m --shuffle" >> $out
  TPB_SUBKEY=$sk timeout 900 python bench.py --quick --workload dam_break_3d_100m --steps 4 --warmup 3 --shuffle 2>/dev/null | tail -1 | cut -c1-330 >> $out
done
cat $out
