#!/bin/bash
# 2 GPUs: fused SSP tests with the bias gradients written into the arena (remaining budget of the round).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 130 python -m pytest tests/test_dist_gpu.py -q -x -k "staleness_zero or straggler" 2>&1 | tail -6 > gpurun_out/r2_call28.log
cat gpurun_out/r2_call28.log
