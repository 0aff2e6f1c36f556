#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> '<command>'  — retries while the pod answers "busy / transient" (nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|nothing was charged\|no box or slot"; then
    sleep 45; continue
  fi
  echo "$out"; exit 0
done
echo "gave up: $out"
