#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/models.log
echo "=== first-layer tests" > $L
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -p no:cacheprovider -k "first_layer or transform" 2>&1 | tail -5 >> $L
for m in vgg16; do
  echo "=== $m" >> $L
  timeout 600 python bench.py --model $m --steps 6 --warmup 3 --no-e2e 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 >> $L
done
cut -c1-300 $L
