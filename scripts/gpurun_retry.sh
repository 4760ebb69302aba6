#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun args]
# Retries while the pod has no free GPU slot (gpurun exit code 3: nothing charged).
T=$1; shift; CMD=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
