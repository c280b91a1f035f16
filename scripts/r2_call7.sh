#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call7.log
echo "=== lrn within debug" > $L
timeout 300 python scripts/debug/lrn_within_dbg.py >> $L 2>&1
echo "=== pair/mcast tests" >> $L
timeout 900 python -m pytest tests/test_pair_cta_gpu.py -q -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== layer kernel tests" >> $L
timeout 600 python -m pytest tests/test_layer_kernels_gpu.py -q -p no:cacheprovider 2>&1 | grep -E "Error|passed|failed|what =" | head -20 >> $L
echo "=== ops+engine tests" >> $L
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -8 >> $L
for mc in 1 2 4; do
  echo "=== conv bench mcast $mc" >> $L
  PSD_MCAST=$mc PSD_PAIR=0 timeout 300 python benchmarks/conv_bench.py conv1,conv3,conv4,conv5 >> $L 2>&1
done
echo "=== bench alexnet" >> $L
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-220 >> $L
echo "=== bench googlenet / vgg16" >> $L
timeout 600 python bench.py --model googlenet --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-220 >> $L
timeout 600 python bench.py --model vgg16 --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 | cut -c1-220 >> $L
tail -100 $L
