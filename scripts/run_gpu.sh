#!/bin/bash
# usage: run_gpu.sh <logname> <timeout> [--gpus N] -- <command>   (retries while the pod is busy)
name=$1; shift; to=$1; shift
extra=""
if [ "$1" == "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > gpurun_out/$name.log 2>&1
  if grep -q "status=transient\|exit code 3\|rc=3" gpurun_out/$name.log; then sleep 60; continue; fi
  break
done
