#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call8.log
echo "=== lrn within debug" > $L
timeout 300 python scripts/debug/lrn_within_dbg.py >> $L 2>&1
echo "=== all 1-GPU tests" >> $L
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12 >> $L
echo "=== bench alexnet" >> $L
timeout 600 python bench.py --steps 100 --warmup 10 2>&1 | tail -1 | cut -c1-400 >> $L
tail -60 $L
