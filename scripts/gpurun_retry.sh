#!/bin/bash
# usage: gpurun_retry.sh <timeout> <gpus> <command...>   -- retries while the pod answers busy (nothing is charged for those)
T=$1; G=$2; shift 2
for i in $(seq 1 20); do
  if [ "$G" = "1" ]; then out=$(gpurun --timeout $T -- "$@" 2>&1); else out=$(gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|rc=3\|busy"; then sleep 150; else break; fi
done
