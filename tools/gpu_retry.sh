#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'   — retries while the pod is busy (exit 3), at most 20 times
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
