#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpurun args...>   — retries while the pod answers "busy / transient" (exit code 3)
for i in $(seq 1 20); do
  gpurun "$@"; rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[retry $i] pod busy, sleeping 150 s"; sleep 150
done
exit 3
