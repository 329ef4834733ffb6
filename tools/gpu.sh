#!/bin/bash
# tools/gpu.sh TIMEOUT 'command' -- gpurun with retries while the pod is busy (exit code 3 = nothing charged)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
