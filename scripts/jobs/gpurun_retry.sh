#!/bin/bash
# usage: gpurun_retry.sh <timeout> <job script> <log>: retries while the pod answers "busy" (exit 3)
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "bash $2" > "$3" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 100
done
exit 3
