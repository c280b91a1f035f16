#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/all.log
echo "=== gemm+ops+engine tests" > $L
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -q -p no:cacheprovider 2>&1 | tail -30 >> $L
echo "=== gemm bench" >> $L
timeout 300 python benchmarks/gemm_bench.py >> $L 2>&1
echo "=== conv bench" >> $L
timeout 600 python benchmarks/conv_bench.py >> $L 2>&1
echo "=== bench sm100 alexnet (graph)" >> $L
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 >> $L
echo "=== bench sm100 alexnet (eager)" >> $L
timeout 900 python bench.py --steps 20 --warmup 5 --graph 0 --no-e2e 2>&1 | tail -1 >> $L
echo "=== bench sm100 googlenet" >> $L
timeout 900 python bench.py --model googlenet --steps 20 --warmup 5 2>&1 | tail -2 >> $L
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --graph 0 > gpurun_out/prof_bench.log 2>&1
tail -70 $L
