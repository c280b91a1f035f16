#!/bin/bash
# usage: scripts/gpu_dist.sh <ngpu>
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/dist$N.log
nvidia-smi topo -m > $L 2>&1
echo "=== dist tests" >> $L
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_dist_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -40 >> $L
echo "=== bench sm100 fused N=$N" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -12 >> $L
echo "=== bench torch nccl N=$N" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --engine torch --comm nccl --svb 0 2>&1 | tail -4 >> $L
tail -100 $L
