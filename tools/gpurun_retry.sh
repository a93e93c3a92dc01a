#!/bin/bash
# developer tool: gpurun with retries while the pod answers "busy" (exit code 3)
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT 'command'
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
