#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy" (exit code 3)
log="$1"; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
