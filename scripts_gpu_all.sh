#!/bin/bash
cd "$(dirname "$0")"
mkdir -p gpurun_out
L=gpurun_out/all.log
echo "=== gemm+ops+engine tests" > $L
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -30 >> $L
echo "=== bench sm100 alexnet" >> $L
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 >> $L
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 560 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e > gpurun_out/prof_bench.log 2>&1
tail -60 $L
