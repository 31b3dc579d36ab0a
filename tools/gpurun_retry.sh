#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3, nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>' [extra gpurun args]
t=$1; shift
cmd=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$t" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
