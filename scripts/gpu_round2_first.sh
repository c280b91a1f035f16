#!/bin/bash
# First GPU call of the next round (ROADMAP.md, "First GPU calls"): re-validate everything that changed while no GPU was
# available.  One B200, ≈ 8 minutes.      gpurun --timeout 900 -- bash scripts/gpu_round2_first.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/round2_first.log
echo "=== full GPU suite (XPASS = tests added without GPU access that pass on hardware)" > $L
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider -rxX 2>&1 | tail -25 >> $L
echo "=== bench alexnet (expect 80.0 k img/s +- 1 %, e2e >= 97 % of value)" >> $L
timeout 120 python bench.py --steps 20 --warmup 5 2>&1 | grep -E '^\{|Error|error' | tail -2 >> $L
echo "=== experimental paired-CTA GEMM (a trap / timeout here = protocol bug, not a framework failure)" >> $L
POSEIDON_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_pair_gemm_gpu.py -q -x -s -p no:cacheprovider 2>&1 | tail -8 >> $L
echo "=== GoogLeNet, default vs channel-padded K" >> $L
timeout 90 python bench.py --model googlenet --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | cut -c1-160 >> $L
POSEIDON_PAD_K=1 timeout 90 python bench.py --model googlenet --steps 20 --warmup 5 --no-e2e 2>&1 | grep -E '^\{' | cut -c1-160 >> $L
cut -c1-300 $L
