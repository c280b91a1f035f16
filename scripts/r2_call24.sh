#!/bin/bash
# Final 1-GPU validation of the tree: every GPU test, smoke(), the four model benches (default flags of the driver for
# AlexNet), the reference arm's answer, per-kernel lists of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call24.log
echo "=== pytest -m gpu" > $L
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 >> $L
echo "=== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
echo "=== bench.py (driver form)" >> $L
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 10 2>&1 | tail -1 | cut -c1-3000 >> $L
timeout 100 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 >> $L
for M in googlenet vgg16 caffenet; do
  echo "=== $M" >> $L
  timeout 600 python bench.py --model $M --steps 150 --warmup 10 2>&1 | tail -1 | cut -c1-3000 >> $L
done
echo "=== kernel lists" >> $L
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --kernel-list gpurun_out/r2_kernels_alexnet_1gpu_final.txt 2>&1 | tail -1 | cut -c1-200 >> $L
timeout 300 python bench.py --model googlenet --steps 20 --warmup 5 --no-e2e --kernel-list gpurun_out/r2_kernels_googlenet_1gpu_final.txt 2>&1 | tail -1 | cut -c1-200 >> $L
echo "=== pool / lrn micro (conv_bench has none): pool tests timing via pytest durations" >> $L
timeout 300 python -m pytest tests/test_ops_gpu.py -q -k pool 2>&1 | tail -2 >> $L
cat $L | cut -c1-500
