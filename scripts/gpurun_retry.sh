#!/bin/bash
# builder tool: gpurun with retries while the pod answers "transient" (nothing charged); usage: gpurun_retry.sh <timeout> <command>
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  echo "$out" | tail -40
  if ! echo "$out" | grep -q "status=transient\|status=refused"; then exit 0; fi
  sleep 90
done
