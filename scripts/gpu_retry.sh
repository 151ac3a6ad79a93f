#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout_s> [--gpus N] '<command>'  -- retries while the pod answers busy (exit 3)
t=$1; shift
opts=()
if [ "$1" = "--gpus" ]; then opts=(--gpus "$2"); shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "${opts[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
