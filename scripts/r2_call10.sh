#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call10.log
echo "=== E1: plain GEMM with conv3's extent (M 43264, N 384, K 2304), pair and single" > $L
PSD_PAIR=1 timeout 300 python benchmarks/gemm_bench.py 43264 384 2304 20 2>&1 | head -8 >> $L
PSD_PAIR=0 timeout 300 python benchmarks/gemm_bench.py 43264 384 2304 20 2>&1 | head -8 >> $L
echo "=== E2/E5: conv3 variants, single CTA" >> $L
PSD_PAIR=0 timeout 300 python benchmarks/conv_bench.py conv3,conv3_nopad,conv3_as1x1 20 >> $L 2>&1
echo "--- pair" >> $L
PSD_PAIR=1 timeout 300 python benchmarks/conv_bench.py conv3,conv3_nopad,conv3_as1x1 20 >> $L 2>&1
echo "--- ring depth 2 / 3 (single CTA; default 4 at BN 192/256)" >> $L
PSD_MAX_STAGES=2 PSD_PAIR=0 timeout 300 python benchmarks/conv_bench.py conv3 20 >> $L 2>&1
PSD_MAX_STAGES=3 PSD_PAIR=0 timeout 300 python benchmarks/conv_bench.py conv3 20 >> $L 2>&1
echo "--- gather (no im2col TMA)" >> $L
PSD_CONV_IM2COL=0 PSD_PAIR=0 timeout 300 python benchmarks/conv_bench.py conv3,conv3_as1x1 20 >> $L 2>&1
echo "=== launch list alexnet eager" >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 330 --csv --log-file gpurun_out/r2_launches_alexnet.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --graph 0 > gpurun_out/r2_prof_bench.log 2>&1
timeout 60 python benchmarks/launch_summary.py gpurun_out/r2_launches_alexnet.csv 30 >> $L 2>&1
echo "=== kernel lists (torch.profiler)" >> $L
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --kernel-list gpurun_out/r2_kernels_alexnet_1gpu.txt 2>&1 | tail -1 | cut -c1-100 >> $L
timeout 600 python bench.py --model googlenet --steps 5 --warmup 3 --no-e2e --kernel-list gpurun_out/r2_kernels_googlenet_1gpu.txt 2>&1 | tail -1 | cut -c1-100 >> $L
tail -150 $L
