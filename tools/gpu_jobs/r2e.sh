#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "summation") 2>&1 | tail -5
echo "-- tiles"; timeout 200 python bench.py --steps 20 --warmup 3 --quick --summation 2>&1 | tail -1 | cut -c1-400
