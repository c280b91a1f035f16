#!/bin/bash
# 2 GPUs: the whole multi-GPU test file with lanes + deferred weight-gradient joins in the tree, then GoogLeNet (8 lanes).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call22.log
: > $L
timeout 1500 python -m pytest tests/test_dist_gpu.py -q 2>&1 | tail -8 >> $L
echo "=== googlenet 2 GPUs" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --model googlenet --gpus 2 --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 >> $L
cat $L
