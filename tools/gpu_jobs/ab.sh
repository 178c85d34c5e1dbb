#!/bin/bash
# A/B of library variants: tools/gpu_jobs/ab.sh  (variants listed in build_variants/ab_list.txt, "default" = in-tree lib)
mkdir -p gpurun_out
cp trixiparticles.jl_b200/libtpb200.so /tmp/lib_default.so
for v in $(cat build_variants/ab_list.txt); do
  if [ "$v" = "default" ]; then cp /tmp/lib_default.so trixiparticles.jl_b200/libtpb200.so; else cp build_variants/lib_$v.so trixiparticles.jl_b200/libtpb200.so; fi
  echo "== variant $v"
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "neighbor_sets_bit_exact or kick_dam_break_3d or chunked or stress or medium" 2>&1 | tail -2
  for w in dam_break_3d_1m dam_break_3d_10m; do
    timeout 200 python bench.py --steps 20 --warmup 3 --quick --workload $w 2>&1 | tail -1 | cut -c1-420
  done
done
cp /tmp/lib_default.so trixiparticles.jl_b200/libtpb200.so
