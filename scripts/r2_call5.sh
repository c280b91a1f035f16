#!/bin/bash
# ncu captures: conv3 (fprop + dgrad + wgrad) in pair and single mode, plain GEMM pair mode.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call5.log
echo "=== ncu conv3 pair" > $L
PSD_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 6 -c 3 -f -o gpurun_out/r2_conv3_pair \
   python benchmarks/conv_bench.py conv3 2 >> $L 2>&1
echo "=== ncu conv3 single" >> $L
PSD_PAIR=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 6 -c 3 -f -o gpurun_out/r2_conv3_single \
   python benchmarks/conv_bench.py conv3 2 >> $L 2>&1
echo "=== ncu gemm 4096^3 pair" >> $L
PSD_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -f -o gpurun_out/r2_gemm_pair \
   python benchmarks/gemm_bench.py 4096 4096 4096 1 >> $L 2>&1
echo "=== layer kernel tests" >> $L
timeout 600 python -m pytest tests/test_layer_kernels_gpu.py -q -x -p no:cacheprovider 2>&1 | grep -E "Error|error|passed|failed|what" | head -20 >> $L
echo "=== gemm f32 (bulk-store epilogue) tests + fc wgrad timing" >> $L
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_pair_cta_gpu.py -q -p no:cacheprovider -k "f32 or wgrad or sfb" 2>&1 | tail -5 >> $L
timeout 300 python benchmarks/sgd_bench.py >> $L 2>&1
tail -80 $L
