#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout_s> <job script> [max tries]
# Retries a gpurun call while the pod answers busy / transient (nothing charged then).
gpus=$1; tmo=$2; job=$3; tries=${4:-20}
for i in $(seq 1 $tries); do
  if [ "$gpus" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $tmo -- "bash $job" > /tmp/gpurun_retry_$$.log 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus $gpus --timeout $tmo -- "bash $job" > /tmp/gpurun_retry_$$.log 2>&1
  fi
  rc=$?
  if grep -q "status=transient\|status=busy\|status=refused" /tmp/gpurun_retry_$$.log && [ $rc -ne 0 -o -n "$(grep -l 'nothing was charged' /tmp/gpurun_retry_$$.log)" ]; then
    echo "[retry $i] busy: $(grep -o 'status=[a-z]*' /tmp/gpurun_retry_$$.log | head -1)"; sleep 120; continue
  fi
  tail -60 /tmp/gpurun_retry_$$.log; exit $rc
done
echo "gave up after $tries tries"; tail -5 /tmp/gpurun_retry_$$.log; exit 3
