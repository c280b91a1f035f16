#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/engine.log
echo "=== engine tests" > $L
timeout 900 python -m pytest tests/test_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -60 >> $L
echo "=== bench sm100 alexnet" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -30 >> $L
echo "=== bench torch(bf16 vendor) alexnet" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 --engine torch 2>&1 | tail -8 >> $L
tail -120 $L
