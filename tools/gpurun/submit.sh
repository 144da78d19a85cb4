#!/bin/bash
# submit.sh <log> <timeout_s> <gpus> <command...>: gpurun with retries while the pod answers "busy / draining" (exit 3 or
# status=transient: nothing charged).  Run in the background; poll the log.
log=$1; to=$2; gpus=$3; shift 3
for try in $(seq 1 30); do
  if [ "$gpus" -gt 1 ]; then /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > "$log" 2>&1; else /usr/local/graft/bin/gpurun --timeout $to -- "$@" > "$log" 2>&1; fi
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "[submit] finished rc=$rc tries=$try" >> "$log"
