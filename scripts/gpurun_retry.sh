#!/bin/bash
# gpurun_retry.sh LOGFILE [gpurun args...]: retries while the pod answers "transient" (nothing charged) or while an
# earlier call of this repo is still in flight ("refused")
LOG=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if grep -q "status=transient" "$LOG"; then sleep 90;
  elif grep -q "status=refused" "$LOG"; then sleep 45;
  else break; fi
done
