#!/bin/bash
# compute-sanitizer over the round-2 kernels
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
(timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fsi.py -m gpu -x -q -k "(test_kick_dam_break_3d and not medium) or continuity_density_wall or several_wall or adaptive_cole_device or test_fsi_kick_matches_oracle or summation_density") > gpurun_out/r2v_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2v_memcheck.log
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2v_memcheck.log | tail -5
(timeout 900 $S --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2v_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2v_racecheck.log
grep -E "smoke ok|RACECHECK SUMMARY" gpurun_out/r2v_racecheck.log | tail -3
(timeout 600 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_slabs.py -m gpu -x -q -k "adaptive_cole_across or slab_ranks") > gpurun_out/r2v_memcheck_slabs.log 2>&1; echo "memcheck slabs rc=$?" | tee -a gpurun_out/r2v_memcheck_slabs.log
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2v_memcheck_slabs.log | tail -3
