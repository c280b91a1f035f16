#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/last.log
echo "=== conv + engine tests" > $L
timeout 100 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -x -p no:cacheprovider -k "conv or engine or parity or alexnet or googlenet" 2>&1 | tail -6 >> $L
echo "=== bench alexnet" >> $L
timeout 60 python bench.py --steps 10 --warmup 3 --no-e2e 2>&1 | grep -E '^\{|Error|error' | tail -2 >> $L
cut -c1-260 $L
