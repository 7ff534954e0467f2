#!/bin/bash
# usage: scratch/gpu.sh <timeout_s> '<command>'   -- retries while the pod is busy (rc 3)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
