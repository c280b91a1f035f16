#!/bin/bash
# 2 GPUs: NUMA pinning of the rank processes (utils/affinity.py) on / off, AlexNet with the end-to-end leg.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call25.log
: > $L
nvidia-smi topo -m 2>/dev/null | head -8 >> $L
lscpu | grep -i "numa\|socket\|^CPU(s)" >> $L
P=29571
for MODE in center off center off; do
  echo "=== POSEIDON_NUMA=$MODE" >> $L
  POSEIDON_NUMA=$MODE timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 100 --warmup 10 --no-exposed-comm 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'ms')" >> $L 2>&1
  P=$((P+1))
done
cat $L
