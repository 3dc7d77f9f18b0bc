#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged)
for i in $(seq 1 ${GPURUN_TRIES:-15}); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 100
done
exit 3
