#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3: nothing charged).  usage: tools/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] busy, attempt $attempt; sleeping 90 s" >&2
  sleep 90
done
exit 3
