#!/bin/bash
# Usage: scripts/gpu_retry.sh <log> [gpurun args...] -- <command>
# Retries while the pod answers "busy / transient" (exit code 3); stops on the first real verdict.
# Honors /tmp/psd_build.lock: never starts a push while an extension build is in progress.
log=$1; shift
for i in $(seq 1 60); do
  while [ -f /tmp/psd_build.lock ]; do sleep 3; done
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[gpu_retry] verdict rc=$rc after $i tries" >> "$log"; exit $rc; fi
  sleep 60
done
echo "[gpu_retry] gave up" >> "$log"
