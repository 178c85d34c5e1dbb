#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slabs.py -m gpu -q > gpurun_out/r3k_slabs.log 2>&1; tail -3 gpurun_out/r3k_slabs.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3k_bench_2gpu.json 2> gpurun_out/r3k_bench_2gpu.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r3k_bench_2gpu.json; tail -3 gpurun_out/r3k_bench_2gpu.err
