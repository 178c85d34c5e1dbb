#!/bin/bash
# final artefacts with the final library: launch list, full captures of the two tile kernels, reference arm
mkdir -p gpurun_out
T=r3t
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_bench_$T.log 2>&1; echo "ncu launches rc=$?"
cap() {  # name, kernel regex, bench args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o gpurun_out/prof_${name}_$T python bench.py --steps 2 --warmup 3 "$@" > gpurun_out/ncu_full_${name}_$T.log 2>&1; echo "ncu $name rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_${name}_$T.ncu-rep > gpurun_out/ncu_${name}_$T.txt 2>&1
  python tools/ncu_source.py gpurun_out/prof_${name}_$T.ncu-rep 40 > gpurun_out/ncu_${name}_source_top40_$T.txt 2>&1
  rm -f gpurun_out/prof_${name}_$T.ncu-rep
}
cap interact k_interact_tiles --no-cpu-baseline --no-variants
cap interact_shuffled k_interact_tiles --quick --shuffle
cap adami k_adami_tiles --no-cpu-baseline --no-variants
du -sh gpurun_out
