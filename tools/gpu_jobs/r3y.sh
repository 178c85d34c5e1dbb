#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3y_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r3y_smoke.log
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r3y_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3y_pytest_gpu.log; tail -5 gpurun_out/r3y_pytest_gpu.log
