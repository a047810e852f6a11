#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> '<command>'   - retries while the pod answers busy (rc 3 / transient)
t=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|no box\|rc=3"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
