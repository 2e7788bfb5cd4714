#!/bin/bash
# usage: profiles/gpu_retry.sh <logfile> <timeout-seconds> '<command>'   -- retries while the pod answers busy (exit 3)
log=$1; to=$2; cmd=$3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$cmd" > "$log" 2>&1
  rc=$?
  if ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 45
done
exit 3
