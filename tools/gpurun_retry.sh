#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
# usage: tools/gpurun_retry.sh <timeout_s> '<command>' [--gpus N]
T=$1; shift; CMD=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then
    echo "[retry $i] busy, sleeping" >&2; sleep 150; continue
  fi
  cat /tmp/gpurun_last.txt; exit $rc
done
cat /tmp/gpurun_last.txt; exit 3
