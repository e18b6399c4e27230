#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...> : retries while the pod answers "busy" (exit 3 / transient)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|rc=3\|no box or slot" $LOG; then sleep 90; continue; fi
  break
done
