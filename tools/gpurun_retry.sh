#!/usr/bin/env bash
# tools/gpurun_retry.sh [gpurun args...] -- '<command>': calls gpurun, retrying every 2 minutes while the pod answers "busy" (exit 3)
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
