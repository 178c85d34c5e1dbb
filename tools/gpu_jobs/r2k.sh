#!/bin/bash
# round 2 (k): 8-GPU box: slab tests (2 GPUs), bench at N = 8 and N = 4 (config-4 sizes, parity check inside)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
(time timeout 500 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu) > gpurun_out/r2k2_slabtests.log 2>&1
tail -4 gpurun_out/r2k2_slabtests.log
for n in 8 4; do
  (time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3) > gpurun_out/r2k2_bench$n.json 2> gpurun_out/r2k2_bench$n.err
  echo "bench $n rc=$?"; tail -c 600 gpurun_out/r2k2_bench$n.err; head -c 900 gpurun_out/r2k2_bench$n.json; echo
done
