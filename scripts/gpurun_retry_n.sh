#!/bin/bash
# usage: scripts/gpurun_retry_n.sh <gpus> <timeout-seconds> '<command>'  -- multi-GPU variant of gpurun_retry.sh
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus "$1" --timeout "$2" -- "$3"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
