#!/bin/bash
# tools/gpurun_retry.sh <timeout_s> '<command>' — gpurun, retried while the pod answers "busy" (exit 3: nothing charged)
t=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
