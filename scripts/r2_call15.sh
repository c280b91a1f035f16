#!/bin/bash
# bf16 bulk-store epilogue: numerics, then A/B of the step time (POSEIDON_BULK_EPI=1: fp32 outputs only = before).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call15.log
: > $L
echo "=== numerics" >> $L
timeout 900 python -m pytest tests/test_pair_cta_gpu.py -q -x 2>&1 | tail -5 >> $L
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_layer_kernels_gpu.py -q -x 2>&1 | tail -5 >> $L
echo "=== GEMM with conv3's extent" >> $L
timeout 300 python benchmarks/gemm_bench.py 43264 384 2304 20 2>&1 | tail -14 >> $L
POSEIDON_BULK_EPI=1 timeout 300 python benchmarks/gemm_bench.py 43264 384 2304 20 2>&1 | grep "bf16 out" | sed 's/^/walk: /' >> $L
echo "=== GEMM 8192^3" >> $L
timeout 300 python benchmarks/gemm_bench.py 8192 8192 8192 5 2>&1 | grep "BN=256\|cuBLAS" >> $L
echo "=== conv bench" >> $L
timeout 300 python benchmarks/conv_bench.py conv2,conv3,conv4,conv5 20 2>&1 | tail -6 >> $L
POSEIDON_BULK_EPI=1 timeout 300 python benchmarks/conv_bench.py conv2,conv3,conv4,conv5 20 2>&1 | tail -4 | sed 's/^/walk: /' >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
for MODEL in "" "--model googlenet" "--model vgg16"; do
  echo "=== A/B $MODEL" >> $L
  run "default (bf16 + fp32 bulk epilogues)" X=1
  run "bf16 through the per-warp walk" POSEIDON_BULK_EPI=1
  run "default (repeat)" X=1
  run "walk (repeat)" POSEIDON_BULK_EPI=1
done
cat $L
