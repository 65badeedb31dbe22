#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit 3 / status=transient).  Usage:
#   tools/gpurun_retry.sh [--gpus N] <timeout> '<command>'
G=""
if [ "$1" == "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|no box\|retry in a few minutes"; then
    sleep 150
    continue
  fi
  echo "$out"
  exit 0
done
echo "gpurun_retry: gave up"; exit 3
