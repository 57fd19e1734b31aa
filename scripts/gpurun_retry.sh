#!/bin/bash
# gpurun_retry.sh LOGFILE [gpurun args...]: retries while the pod answers "transient" (nothing charged)
LOG=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 90; else break; fi
done
