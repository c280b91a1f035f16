#!/bin/bash
# 2 GPUs: the tests the -x of call 17 cut off (lanes under the fused backend, eager + graph) and the SSP straggler test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call18.log
: > $L
timeout 1200 python -m pytest tests/test_dist_gpu.py -q -k "straggler or lanes" 2>&1 | tail -12 >> $L
cat $L
