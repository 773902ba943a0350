#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun flags]
# Retries while the pod answers "busy" (exit 3: nothing charged); any other exit code ends the loop.
t=$1; shift
cmd=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "$@" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
