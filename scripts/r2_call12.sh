#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call12.log
echo "=== all 1-GPU tests" > $L
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== bench alexnet" >> $L
timeout 600 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | cut -c1-2500 >> $L
echo "=== bench googlenet" >> $L
timeout 600 python bench.py --model googlenet --steps 100 --warmup 10 2>&1 | tail -1 | cut -c1-400 >> $L
echo "=== bench vgg16 / caffenet" >> $L
timeout 600 python bench.py --model vgg16 --steps 30 --warmup 5 --no-e2e 2>&1 | tail -1 | cut -c1-300 >> $L
timeout 600 python bench.py --model caffenet --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-300 >> $L
echo "=== GEMM with conv3's extent incl. cuBLAS" >> $L
PSD_PAIR=1 timeout 300 python benchmarks/gemm_bench.py 43264 384 2304 20 2>&1 | sed -n 1,11p >> $L
echo "=== kernel list googlenet" >> $L
timeout 600 python bench.py --model googlenet --steps 5 --warmup 3 --no-e2e --kernel-list gpurun_out/r2_kernels_googlenet_1gpu.txt 2>&1 | tail -1 | cut -c1-100 >> $L
tail -60 $L
