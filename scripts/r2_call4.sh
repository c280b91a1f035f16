#!/bin/bash
# 2 GPUs: multi-GPU correctness debt + the per-bucket all-reduce launch + SFB variants + vendor arm + exposed comm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call4.log
nvidia-smi topo -m > $L 2>&1
echo "=== dist gpu tests (2 GPUs)" >> $L
timeout 1500 python -m pytest tests/test_dist_gpu.py -q -p no:cacheprovider 2>&1 | tail -25 >> $L
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
echo "=== bench alexnet 2 GPUs (fused SFB epilogue)" >> $L
NCCL_DEBUG=INFO timeout 900 $B --steps 100 --warmup 10 2>&1 | grep -E "NVLS|^\{|symmetric arena|Error|error" | head -12 >> $L
echo "=== bench alexnet 2 GPUs (two-pass SFB)" >> $L
POSEIDON_SFB_FUSED_SGD=0 timeout 900 $B --steps 100 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -1 | cut -c1-900 >> $L
echo "=== bench alexnet 2 GPUs (no multimem: P2P)" >> $L
POSEIDON_MULTIMEM=0 timeout 900 $B --steps 100 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -1 | cut -c1-900 >> $L
echo "=== bench googlenet 2 GPUs" >> $L
timeout 900 $B --model googlenet --steps 50 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-900 >> $L
echo "=== vendor arm 2 GPUs (torch engine + NCCL, CUDA graph)" >> $L
timeout 900 $B --engine torch --graph 1 --steps 50 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -2 | cut -c1-900 >> $L
timeout 900 $B --engine torch --graph 1 --model googlenet --steps 50 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -2 | cut -c1-900 >> $L
echo "=== caffenet SSP staleness 1, 2 GPUs (fused SSP kernels, CUDA graph)" >> $L
timeout 900 $B --model caffenet --staleness 1 --steps 50 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -1 | cut -c1-600 >> $L
echo "=== bench alexnet 1 GPU (same box)" >> $L
timeout 600 python bench.py --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 | cut -c1-900 >> $L
tail -120 $L
