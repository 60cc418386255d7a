#!/bin/bash
# gpurun with retries while the pod answers "busy / transient" (nothing charged in that case).
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [gpus]
T=$1; CMD=$2; G=${3:-1}
for attempt in $(seq 1 40); do
  if [ "$G" = "1" ]; then
    OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$CMD" 2>&1)
  else
    OUT=$(/usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$CMD" 2>&1)
  fi
  if echo "$OUT" | grep -q "status=transient\|nothing was charged — retry\|exit code 3"; then
    sleep 90
    continue
  fi
  echo "$OUT"
  exit 0
done
echo "$OUT"
echo "gave up after 40 attempts"
