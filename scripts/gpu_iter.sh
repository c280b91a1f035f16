#!/bin/bash
# quick iteration: numerics tests + micro benches + one AlexNet bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/iter.log
echo "=== tests" > $L
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -25 >> $L
echo "=== sgd bench" >> $L
timeout 300 python benchmarks/sgd_bench.py >> $L 2>&1
echo "=== conv bench (im2col TMA on)" >> $L
timeout 300 python benchmarks/conv_bench.py conv2,conv3,conv4,conv5 5 >> $L 2>&1
echo "=== conv bench (im2col TMA off)" >> $L
PSD_CONV_IM2COL=0 timeout 300 python benchmarks/conv_bench.py conv2,conv3,conv4,conv5 5 >> $L 2>&1
echo "=== bench alexnet (graph)" >> $L
timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 >> $L
echo "=== bench googlenet (graph)" >> $L
timeout 900 python bench.py --model googlenet --steps 20 --warmup 5 2>&1 | tail -1 >> $L
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --graph 0 > gpurun_out/prof_bench.log 2>&1
tail -40 $L
