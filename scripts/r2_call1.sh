#!/bin/bash
# Round-2 call 1: validation debt on 1 GPU + the never-run paired-CTA experiment + baseline bench numbers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call1.log
echo "=== pytest -m gpu (1 GPU)" > $L
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -25 >> $L
echo "=== pair gemm (experimental, cta_group::2)" >> $L
POSEIDON_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_pair_gemm_gpu.py -q -p no:cacheprovider 2>&1 | tail -40 >> $L
echo "=== bench alexnet" >> $L
timeout 600 python bench.py --steps 200 --warmup 10 2>&1 | tail -2 >> $L
echo "=== bench googlenet" >> $L
timeout 600 python bench.py --model googlenet --steps 100 --warmup 10 2>&1 | tail -2 >> $L
echo "=== gemm bench" >> $L
timeout 300 python benchmarks/gemm_bench.py >> $L 2>&1
nvidia-smi -q | grep -i -E "product name|fabric|nvlink" | head >> $L
tail -120 $L
