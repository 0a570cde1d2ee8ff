#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun args ...] -- 'command'   (retries while the pod answers "no box right now")
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
