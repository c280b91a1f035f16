#!/bin/bash
# Deferred join of the weight-gradient side streams: numerics, then step times against the per-layer join.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call20.log
: > $L
echo "=== numerics" >> $L
timeout 900 python -m pytest tests/test_lanes.py tests/test_engine_gpu.py -q -x 2>&1 | tail -6 >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
MODEL=""
echo "=== alexnet" >> $L
run "deferred join (default)" X=1
run "join per layer" POSEIDON_WGRAD_DEFER=0
run "deferred join (repeat)" X=1
run "join per layer (repeat)" POSEIDON_WGRAD_DEFER=0
MODEL="--model googlenet"
echo "=== $MODEL" >> $L
run "deferred, 4 lanes" X=1
run "per layer, 4 lanes" POSEIDON_WGRAD_DEFER=0
run "deferred, 8 lanes" POSEIDON_LANES=8
MODEL="--model vgg16"
echo "=== vgg16" >> $L
run "deferred" X=1
run "per layer" POSEIDON_WGRAD_DEFER=0
MODEL="--model caffenet"
echo "=== caffenet" >> $L
run "deferred" X=1
cat $L
