#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/cmp.log
timeout 600 python benchmarks/compare_engines.py googlenet 4 > $L 2>&1
timeout 600 python benchmarks/compare_engines.py alexnet 8 >> $L 2>&1
timeout 900 python -m pytest tests/test_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -15 >> $L
grep -v "cannot open database\|mean_file" $L | tail -60
