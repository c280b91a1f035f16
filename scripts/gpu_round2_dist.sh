#!/bin/bash
# Second GPU call of the next round: 2 GPUs.     gpurun --gpus 2 --timeout 600 -- bash scripts/gpu_round2_dist.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/round2_dist.log
echo "=== 2-GPU tests (fused all-reduce / SFB, NCCL baseline, snapshot + resume for fused / nccl / ssp)" > $L
timeout 400 python -m pytest tests/test_dist_gpu.py -q -p no:cacheprovider 2>&1 | tail -12 >> $L
echo "=== bench alexnet x2 (expect 151 k img/s)" >> $L
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | grep -E '^\{|Error|error' | tail -2 >> $L
cut -c1-300 $L
