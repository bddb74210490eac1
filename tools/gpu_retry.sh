#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'  [extra gpurun args...]
# Retries while gpurun answers "no box / slot free right now" (exit code 3, nothing charged).
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
