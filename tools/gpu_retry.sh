#!/bin/bash
# developer tool: run a gpurun call, retrying while the pod answers busy (exit code 3 = nothing charged)
# usage: tools/gpu_retry.sh <log> <timeout_s> '<command>'
log=$1; to=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
