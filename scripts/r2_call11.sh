#!/bin/bash
# 8 GPUs: AlexNet (DWBP + SFB) with exposed-comm + end-to-end + per-rank kernel list, GoogLeNet, CaffeNet SSP s=1.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call11.log
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8"
echo "=== alexnet 8 GPUs" > $L
NCCL_DEBUG=WARN timeout 600 $B --steps 100 --warmup 10 --kernel-list gpurun_out/r2_kernels_alexnet_8gpu.txt 2>&1 | grep -E "^\{|symmetric arena|rror" | cut -c1-3000 >> $L
echo "=== googlenet 8 GPUs" >> $L
timeout 600 $B --model googlenet --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-3000 >> $L
echo "=== caffenet SSP staleness 1, 8 GPUs" >> $L
timeout 600 $B --model caffenet --staleness 1 --steps 50 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -1 | cut -c1-3000 >> $L
rm -f gpurun_out/r2_kernels_alexnet_8gpu.txt.[2-7]
tail -40 $L | cut -c1-400
