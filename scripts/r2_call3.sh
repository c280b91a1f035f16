#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call3.log
echo "=== pair cta tests" > $L
timeout 900 python -m pytest tests/test_pair_cta_gpu.py -q -x -s -p no:cacheprovider 2>&1 | tail -25 >> $L
echo "=== layer kernel tests" >> $L
timeout 600 python -m pytest tests/test_layer_kernels_gpu.py -q -p no:cacheprovider 2>&1 | tail -25 >> $L
echo "=== gemm+ops+engine tests" >> $L
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -25 >> $L
echo "=== gemm bench (pair)" >> $L
PSD_PAIR=1 timeout 300 python benchmarks/gemm_bench.py 2>&1 | head -14 >> $L
echo "=== conv bench pair / single" >> $L
PSD_PAIR=1 timeout 300 python benchmarks/conv_bench.py >> $L 2>&1
PSD_PAIR=0 timeout 300 python benchmarks/conv_bench.py >> $L 2>&1
echo "=== bench alexnet (pair)" >> $L
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 >> $L
echo "=== bench googlenet / vgg16 (pair)" >> $L
timeout 600 python bench.py --model googlenet --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 >> $L
timeout 600 python bench.py --model vgg16 --steps 20 --warmup 5 --no-e2e 2>&1 | tail -1 >> $L
echo "=== vendor arm (torch engine, CUDA graph)" >> $L
timeout 600 python bench.py --engine torch --graph 1 --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 >> $L
timeout 600 python bench.py --engine torch --graph 1 --model googlenet --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 >> $L
tail -200 $L
