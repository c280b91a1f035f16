#!/bin/bash
# Branch-parallel lanes: numerics (eager + graph), then GoogLeNet step time with 4 lanes vs the sequential schedule.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call16.log
: > $L
echo "=== numerics" >> $L
timeout 600 python -m pytest tests/test_lanes.py -q -x 2>&1 | tail -15 >> $L
timeout 600 python -m pytest tests/test_engine_gpu.py -q -x 2>&1 | tail -5 >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
MODEL="--model googlenet"
echo "=== A/B $MODEL" >> $L
run "4 lanes (default)" X=1
run "sequential" POSEIDON_LANES=1
run "4 lanes (repeat)" X=1
run "sequential (repeat)" POSEIDON_LANES=1
run "2 lanes" POSEIDON_LANES=2
run "6 lanes" POSEIDON_LANES=6
MODEL=""
echo "=== alexnet (no forks: plan must stay single-lane)" >> $L
run "default" X=1
echo "=== googlenet eager, 4 lanes vs 1 (CPU launch bound)" >> $L
timeout 300 python bench.py --model googlenet --steps 30 --warmup 5 --no-e2e --graph 0 2>&1 | tail -1 | cut -c1-160 >> $L
POSEIDON_LANES=1 timeout 300 python bench.py --model googlenet --steps 30 --warmup 5 --no-e2e --graph 0 2>&1 | tail -1 | cut -c1-160 >> $L
cat $L
