#!/bin/bash
# Usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers "busy", exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
