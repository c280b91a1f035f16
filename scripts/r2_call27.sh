#!/bin/bash
# 2 GPUs, last call of the round: bias gradients written into the arena by the column-sum kernel on the side stream.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call27.log
: > $L
timeout 175 python -m pytest tests/test_dist_gpu.py -q -x -k "allreduce or sfb_two or lanes or (snapshot and fused)" 2>&1 | tail -6 >> $L
echo "=== googlenet 2 GPUs" >> $L
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --model googlenet --gpus 2 --steps 60 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -1 | cut -c1-260 >> $L
cat $L
