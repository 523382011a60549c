#!/bin/bash
# usage: [GPUS=N] gpurun_retry.sh LOG TIMEOUT CMD...   retries while the pod answers "busy" (exit 3)
LOG=$1; TO=$2; shift 2
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout "$TO" -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then break; fi
  sleep 90
done
echo "gpurun_retry finished rc=$rc" >> "$LOG"
