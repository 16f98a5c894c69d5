#!/bin/bash
# gpurun with retry while the pod is busy (exit code 3 = nothing charged).  usage: tools/gpu.sh [--gpus N] TIMEOUT 'command'
extra=()
if [ "$1" == "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "${extra[@]}" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
