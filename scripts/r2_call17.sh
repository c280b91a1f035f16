#!/bin/bash
# 2 GPUs: all multi-GPU tests (incl. branch-parallel lanes under the fused backend, eager + graph), GoogLeNet / AlexNet benches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call17.log
: > $L
echo "=== dist tests (2 GPUs)" >> $L
timeout 1200 python -m pytest tests/test_dist_gpu.py -q -x 2>&1 | tail -8 >> $L
P=29541
for M in googlenet alexnet; do
  echo "=== $M 2 GPUs" >> $L
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --model $M --gpus 2 --steps 100 --warmup 10 --no-e2e 2>&1 | tail -1 >> $L
  P=$((P+1))
done
echo "=== googlenet 2 GPUs sequential" >> $L
POSEIDON_LANES=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --model googlenet --gpus 2 --steps 100 --warmup 10 --no-e2e --no-exposed-comm 2>&1 | tail -1 | cut -c1-200 >> $L
cat $L
