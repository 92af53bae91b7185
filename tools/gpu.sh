#!/bin/bash
# Developer helper: run a command on the GPU box through gpurun, retrying while the pod answers "busy".
#   tools/gpu.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
