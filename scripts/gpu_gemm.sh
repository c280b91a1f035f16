#!/bin/bash
# run each GEMM test group in its own process (a device-side trap kills the CUDA context)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gemm_tests.log 2>&1
for t in test_gemm_tn_bf16 test_gemm_dgrad test_gemm_wgrad test_sfb_outer_sgd_single test_sfb_outer_multi; do
  echo "=== $t" >> gpurun_out/gemm_tests.log
  timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k $t -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/gemm_tests.log
done
tail -120 gpurun_out/gemm_tests.log
