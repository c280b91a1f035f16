#!/bin/bash
# Pack-free data gradient (conv_dgrad_w): numerics on the GPU, then A/B of the step time against the packed operand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call14.log
: > $L
echo "=== numerics" >> $L
timeout 900 python -m pytest tests/test_pair_cta_gpu.py -q -x -k "dgrad_from_fprop or conv_im2col" 2>&1 | tail -5 >> $L
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -x 2>&1 | tail -5 >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
for MODEL in "" "--model googlenet" "--model vgg16"; do
  echo "=== A/B $MODEL" >> $L
  run "default (dgrad reads fprop weights)" X=1
  run "packed dgrad operand" POSEIDON_DGRAD_PACK=1
  run "default (repeat)" X=1
  run "packed (repeat)" POSEIDON_DGRAD_PACK=1
done
cat $L
