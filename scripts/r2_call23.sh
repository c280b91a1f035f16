#!/bin/bash
# 8 GPUs, final tree: AlexNet (DWBP + SFB, e2e + exposed comm) and GoogLeNet (8 lanes, e2e + exposed comm).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call23.log
echo "=== alexnet 8 GPUs" > $L
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 100 --warmup 10 2>&1 | grep -E "^\{|rror" | cut -c1-3500 >> $L
echo "=== googlenet 8 GPUs" >> $L
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --model googlenet --steps 100 --warmup 10 --kernel-list gpurun_out/r2_kernels_googlenet_8gpu.txt 2>&1 | grep -E "^\{|rror" | cut -c1-3500 >> $L
rm -f gpurun_out/r2_kernels_googlenet_8gpu.txt.[1-7]
cat $L | cut -c1-600
