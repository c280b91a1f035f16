#!/bin/bash
cd "$(dirname "$0")"
mkdir -p gpurun_out
python tools_conv_bench.py > gpurun_out/conv_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -c 12 -o gpurun_out/conv_prof \
   python tools_conv_bench.py conv1,conv2,conv3 1 > gpurun_out/ncu_conv.log 2>&1
cat gpurun_out/conv_bench.log
ls -la gpurun_out/*.ncu-rep
