#!/bin/bash
# compute-sanitizer over the layer kernels and the small GEMM / conv cases (never executed before).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  L=gpurun_out/r2_sanitize_$TOOL.log
  echo "# compute-sanitizer --tool $TOOL  (layer kernels: tests/test_layer_kernels_gpu.py; ops: lrn/pool/transform/softmax/dropout/colsum/fused_update)" > $L
  timeout 420 compute-sanitizer --tool $TOOL --error-exitcode 1 --launch-timeout 300 \
    python -m pytest tests/test_layer_kernels_gpu.py tests/test_ops_gpu.py -q -x -p no:cacheprovider \
    -k "unary or threshold or eltwise or softmax or mvn or lrn or stochastic or pool or transform or dropout or colsum" 2>&1 | tail -15 >> $L
  echo "exit code: $?" >> $L
done
# memcheck over the tcgen05 paths on small shapes (pair / single, bulk-store epilogue, split-K finish)
L=gpurun_out/r2_sanitize_memcheck_gemm.log
echo "# compute-sanitizer --tool memcheck  (tcgen05 GEMM family, small shapes)" > $L
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 1 --launch-timeout 300 \
  python -m pytest tests/test_pair_cta_gpu.py -q -x -p no:cacheprovider -k "kmajor and (256-256-64 or 384-256-512 or 640-96) or mnmajor_f32 and 512" 2>&1 | tail -12 >> $L
echo "exit code: $?" >> $L
tail -n 8 gpurun_out/r2_sanitize_*.log
