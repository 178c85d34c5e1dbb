#!/bin/bash
# final artefacts of the round: tests, default bench (with variants), launch lists at 1 M and 10 M, full captures
# (summarised on the box: the merged gpurun_out/ must stay below 64 MiB)
mkdir -p gpurun_out
T=r2s
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$T.log; tail -4 gpurun_out/pytest_gpu_$T.log
timeout 900 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_$T.json; echo; tail -3 gpurun_out/bench_$T.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_bench_$T.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_10m_$T.csv python bench.py --steps 2 --warmup 1 --quick --workload dam_break_3d_10m > gpurun_out/ncu_bench_10m_$T.log 2>&1; echo "ncu 10m rc=$?"
cap() {  # name, kernel regex, bench args...
  local name=$1 rx=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o gpurun_out/prof_${name}_$T python bench.py --steps 2 --warmup 3 "$@" > gpurun_out/ncu_full_${name}_$T.log 2>&1; echo "ncu $name rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_${name}_$T.ncu-rep > gpurun_out/ncu_${name}_$T.txt 2>&1
  python tools/ncu_source.py gpurun_out/prof_${name}_$T.ncu-rep 40 > gpurun_out/ncu_${name}_source_top40_$T.txt 2>&1
}
cap interact k_interact_tiles --no-cpu-baseline --no-variants
cap interact_f64 k_interact_tiles --quick --eltype f64 --coords f64
cap adami k_adami_tiles --no-cpu-baseline --no-variants
rm -f gpurun_out/prof_interact_f64_$T.ncu-rep gpurun_out/prof_adami_$T.ncu-rep
for w in dam_break_2d dam_break_3d_small; do echo "== $w"; timeout 200 python bench.py --steps 50 --warmup 5 --quick --workload $w 2>&1 | tail -1 | cut -c1-330; done | tee gpurun_out/small_configs_$T.txt
du -sh gpurun_out
