#!/bin/bash
# tools/gpurun_retry.sh <out-file> <gpurun args...>: re-issues a gpurun call while the pod answers "transient" (nothing charged)
OUT=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$OUT" 2>&1
  if grep -q "status=transient\|exit code 3\|no box" "$OUT" && ! grep -q "status=ok" "$OUT"; then
    echo "attempt $attempt transient; retrying in 120 s" >> "$OUT.retries"; sleep 120
  else
    break
  fi
done
