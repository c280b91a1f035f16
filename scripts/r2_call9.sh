#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call9.log
echo "=== tests: layer kernels, ops (strided dgrad), engine, emulator parity" > $L
timeout 1500 python -m pytest tests/test_layer_kernels_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_zz_emulator_parity_gpu.py -q -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== bench alexnet" >> $L
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-200 >> $L
echo "=== bench googlenet (zero-copy concat)" >> $L
timeout 600 python bench.py --model googlenet --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-200 >> $L
echo "=== gemm bench" >> $L
timeout 300 python benchmarks/gemm_bench.py 2>&1 | tail -16 >> $L
echo "=== launch list alexnet eager" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file gpurun_out/r2_launches_alexnet.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --graph 0 > gpurun_out/r2_prof_bench.log 2>&1
timeout 60 python benchmarks/launch_summary.py gpurun_out/r2_launches_alexnet.csv >> $L 2>&1
tail -150 $L
