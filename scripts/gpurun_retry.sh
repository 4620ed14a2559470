#!/bin/bash
# gpurun with retries while the pod answers "transient" / exit 3 (nothing charged in that case)
# usage: scripts/gpurun_retry.sh <tries> <gpurun args...>
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then
    echo "[retry $i] transient"; sleep 60; continue
  fi
  echo "$out"; exit $rc
done
echo "gave up after $tries tries"; exit 3
