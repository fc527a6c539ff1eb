#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged). Usage: gpurun_retry.sh [gpurun args] -- cmd
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
