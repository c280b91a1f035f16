#!/bin/bash
# One GPU: bias column sums and the inner-product optimizer steps on the weight-gradient side stream — numerics, A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call26.log
: > $L
echo "=== numerics" >> $L
timeout 900 python -m pytest tests/test_lanes.py tests/test_engine_gpu.py tests/test_zz_emulator_parity_gpu.py -q -x 2>&1 | tail -4 >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
MODEL=""
echo "=== alexnet" >> $L
run "default (early IP update + bias sums on the side stream)" X=1
run "IP update at the end of the step" POSEIDON_EARLY_IP_UPDATE=0
run "default (repeat)" X=1
run "end of step (repeat)" POSEIDON_EARLY_IP_UPDATE=0
MODEL="--model vgg16"
echo "=== vgg16" >> $L
run "default" X=1
run "end of step" POSEIDON_EARLY_IP_UPDATE=0
MODEL="--model googlenet"
echo "=== googlenet" >> $L
run "default" X=1
MODEL="--model caffenet"
echo "=== caffenet" >> $L
run "default" X=1
cat $L
