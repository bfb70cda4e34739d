#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing is charged for those).
# usage: scripts/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out"
  if ! echo "$out" | grep -q "status=transient\|exit code 3\|rc=3 "; then exit 0; fi
  echo "[gpurun_retry] attempt $attempt was transient; sleeping 60 s" >&2
  sleep 60
done
exit 3
