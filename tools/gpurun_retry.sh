#!/bin/bash
# tools/gpurun_retry.sh <out-file> <gpurun args...>: retries while the pod answers "busy" (exit 3)
out=$1; shift
for k in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
