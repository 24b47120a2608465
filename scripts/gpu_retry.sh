#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers busy)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_last.txt; then cat /tmp/gpurun_last.txt; exit $rc; fi
  sleep 90
done
cat /tmp/gpurun_last.txt
exit 3
