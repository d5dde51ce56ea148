#!/bin/bash
# usage: tools/r2/gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "busy/transient" (exit 3)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> "$LOG"; exit $rc; fi
  sleep 90
done
echo "rc=3 (gave up)" >> "$LOG"
