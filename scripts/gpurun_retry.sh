#!/bin/bash
# gpurun_retry.sh TIMEOUT 'command' : retry while the pod answers "busy / draining" (nothing charged), up to ~40 min
T=$1; shift
for i in $(seq 1 14); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up: pod busy"
