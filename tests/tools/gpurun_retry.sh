#!/bin/bash
# usage: [GPUS=N] gpurun_retry.sh <timeout_s> '<command>'  — retries while the pod answers "busy / transient" (nothing charged)
T=$1; shift
G=${GPUS:-1}
EXTRA=""
if [ "$G" != "1" ]; then EXTRA="--gpus $G"; fi
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" $EXTRA -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|nothing was charged\|no box or slot"; then
    sleep 45; continue
  fi
  echo "$out"; exit 0
done
echo "gave up: $out"
