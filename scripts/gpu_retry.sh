#!/bin/bash
# gpu_retry.sh <timeout> <command...>: gpurun, retried while the pod answers "busy" (nothing is charged for those)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|no box or slot" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
cat /tmp/gpurun_last.log
exit $rc
