#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged).  usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
