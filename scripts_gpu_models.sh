#!/bin/bash
cd "$(dirname "$0")"
mkdir -p gpurun_out
L=gpurun_out/models.log
echo "=== full gpu test suite" > $L
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 >> $L
for m in vgg16 caffenet; do
  echo "=== $m" >> $L
  timeout 600 python bench.py --model $m --steps 6 --warmup 3 --no-e2e 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3 >> $L
done
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
cut -c1-300 $L
