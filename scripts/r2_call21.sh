#!/bin/bash
# Inner-product weight gradients on the side stream (A/B), the fc wgrad micro-benchmark, one ncu capture of it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call21.log
: > $L
echo "=== numerics" >> $L
timeout 900 python -m pytest tests/test_lanes.py tests/test_engine_gpu.py -q -x 2>&1 | tail -4 >> $L
echo "=== fc wgrad micro-benchmark" >> $L
timeout 300 python benchmarks/fc_wgrad_bench.py 20 >> $L 2>&1
POSEIDON_PAIR_CTA=0 timeout 300 python benchmarks/fc_wgrad_bench.py 20 2>&1 | sed 's/^/single CTA: /' >> $L
POSEIDON_BULK_EPI=0 timeout 300 python benchmarks/fc_wgrad_bench.py 20 2>&1 | sed 's/^/walk epilogue: /' >> $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s', 'launches', d.get('gpu_launches'))" >> $L 2>&1; }
MODEL=""
echo "=== alexnet" >> $L
run "IP wgrad on the side stream (default)" X=1
run "IP wgrad on the main stream" POSEIDON_IP_WGRAD_LANE=0
run "default (repeat)" X=1
run "main stream (repeat)" POSEIDON_IP_WGRAD_LANE=0
MODEL="--model googlenet"
echo "=== $MODEL (8 lanes default)" >> $L
run "default" X=1
MODEL="--model vgg16"
echo "=== vgg16" >> $L
run "default" X=1
run "IP on main" POSEIDON_IP_WGRAD_LANE=0
echo "=== ncu fc6 wgrad" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 4 -c 1 -o gpurun_out/r2_fc6_wgrad -f python benchmarks/fc_wgrad_bench.py 3 one > gpurun_out/r2_fc6_ncu.log 2>&1
tail -3 gpurun_out/r2_fc6_ncu.log >> $L
cat $L
