#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/sgd.log
timeout 300 python benchmarks/sgd_bench.py > $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 2 -c 1 -f -o gpurun_out/sgd_prof \
   python benchmarks/sgd_bench.py 1 >> $L 2>&1
tail -20 $L
