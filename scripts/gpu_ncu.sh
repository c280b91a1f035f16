#!/bin/bash
# ncu --set full captures: conv3 fprop/dgrad/wgrad (implicit GEMM) + the bandwidth-bound layer kernels of one AlexNet step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gemm bench" > gpurun_out/ncu.log
timeout 600 python benchmarks/gemm_bench.py >> gpurun_out/ncu.log 2>&1
echo "=== conv bench" >> gpurun_out/ncu.log
timeout 600 python benchmarks/conv_bench.py conv1,conv2,conv3,conv4,conv5 5 >> gpurun_out/ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 4 -c 3 -f -o gpurun_out/conv3_prof \
   python benchmarks/conv_bench.py conv3 1 >> gpurun_out/ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -s 4 -c 3 -f -o gpurun_out/conv2_prof \
   python benchmarks/conv_bench.py conv2 1 >> gpurun_out/ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:pool_bwd|lrn_reg|pool_fwd|transform_kernel|colsum' -c 14 -f -o gpurun_out/layers_prof \
   python bench.py --steps 1 --warmup 1 --no-e2e --graph 0 >> gpurun_out/ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep >> gpurun_out/ncu.log
tail -60 gpurun_out/ncu.log
