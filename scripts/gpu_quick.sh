#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/quick.log
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_ops_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -30 > $L
echo "=== conv bench (cluster 2)" >> $L
timeout 600 python benchmarks/conv_bench.py >> $L 2>&1
echo "=== bench" >> $L
timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-420 >> $L
cat $L
