#!/bin/bash
# round 2 (a): 2-GPU bench at config-4 sizes + slab tests
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_smi.txt; nproc >> gpurun_out/r2a_smi.txt; free -g >> gpurun_out/r2a_smi.txt
nvidia-smi topo -m >> gpurun_out/r2a_smi.txt 2>&1
(time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3) > gpurun_out/r2a_bench2.json 2> gpurun_out/r2a_bench2.err
tail -c 1500 gpurun_out/r2a_bench2.err
(time timeout 400 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu) > gpurun_out/r2a_slabtests.log 2>&1
tail -5 gpurun_out/r2a_slabtests.log
