#!/bin/bash
# final library on the 8-GPU box: bench at N = 8 (config-4 sizes, parity check inside)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8; do
  (time timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3) > gpurun_out/r3q_bench$n.json 2> gpurun_out/r3q_bench$n.err
  echo "bench $n rc=$?"; tail -c 400 gpurun_out/r3q_bench$n.err; head -c 700 gpurun_out/r3q_bench$n.json; echo
done
