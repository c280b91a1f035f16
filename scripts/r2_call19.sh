#!/bin/bash
# Weight gradients on side streams: numerics, then step times (GoogLeNet lanes 4 / 6 / 8 x wgrad lane; AlexNet / VGG).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call19.log
: > $L
echo "=== numerics" >> $L
timeout 900 python -m pytest tests/test_lanes.py tests/test_engine_gpu.py tests/test_ops_gpu.py -q -x 2>&1 | tail -6 >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
MODEL="--model googlenet"
echo "=== $MODEL" >> $L
run "4 lanes + wgrad lane (default)" X=1
run "4 lanes, wgrad on the layer's stream" POSEIDON_WGRAD_LANE=0
run "6 lanes + wgrad lane" POSEIDON_LANES=6
run "8 lanes + wgrad lane" POSEIDON_LANES=8
run "6 lanes, no wgrad lane" POSEIDON_LANES=6 POSEIDON_WGRAD_LANE=0
run "1 lane + wgrad lane" POSEIDON_LANES=1
MODEL=""
echo "=== alexnet" >> $L
run "wgrad lane (default)" X=1
run "no wgrad lane" POSEIDON_WGRAD_LANE=0
run "wgrad lane (repeat)" X=1
run "no wgrad lane (repeat)" POSEIDON_WGRAD_LANE=0
MODEL="--model vgg16"
echo "=== vgg16" >> $L
run "wgrad lane (default)" X=1
run "no wgrad lane" POSEIDON_WGRAD_LANE=0
cat $L
