#!/bin/bash
# usage: scripts/gpu_scale.sh <ngpu> [tests]   — fused-engine benches at N GPUs (+ the dist test-suite when asked)
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/scale$N.log
echo "=== N=$N" > $L
if [ "$2" == "tests" ]; then
  NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_dist_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -15 >> $L
fi
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 10 --warmup 3 "${@:2}" 2>&1 | grep -E '^\{|Error|error' | tail -3 >> $L; }
if [ "$N" == "8" ]; then
  echo "--- alexnet fused (+e2e)" >> $L; run 29521
  echo "--- googlenet fused" >> $L; run 29522 --model googlenet --no-e2e
  echo "--- alexnet torch+nccl baseline" >> $L; run 29524 --engine torch --comm nccl --svb 0 --no-e2e
  echo "--- vgg16 fused" >> $L; run 29526 --model vgg16 --no-e2e
  echo "--- caffenet ssp staleness 1" >> $L; run 29527 --model caffenet --staleness 1 --no-e2e
else
  echo "--- alexnet fused" >> $L; run 29521 --no-e2e
  echo "--- googlenet fused" >> $L; run 29522 --model googlenet --no-e2e
fi
cut -c1-420 $L
