#!/bin/bash
# Resubmits one gpurun command while the pod answers "transient" (nothing charged); dev tool.
#   scripts/gpurun_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 20); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$OUT" | tail -60
  if ! echo "$OUT" | grep -q "status=transient"; then exit 0; fi
  echo "[retry $i] transient, sleeping 90 s"
  sleep 90
done
exit 3
