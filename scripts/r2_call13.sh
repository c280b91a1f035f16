#!/bin/bash
# A/B on one box: AlexNet / VGG-16 step time under the schedule switches; then the sanitizer runs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_call13.log
: > $L
run() { echo "--- $1" >> $L; shift; env "$@" timeout 600 python bench.py --steps 150 --warmup 10 --no-e2e $MODEL 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['ms_per_step'],4), 'ms', round(d['value']), 'img/s')" >> $L 2>&1; }
for MODEL in "" "--model vgg16"; do
  echo "=== A/B $MODEL" >> $L
  run "default" X=1
  run "default (repeat)" X=1
  run "multi-update off" POSEIDON_MULTI_UPDATE=0
  run "conv pair on" POSEIDON_CONV_PAIR=1
  run "bulk epilogue off" POSEIDON_BULK_EPI=0
  run "GEMM pair off" POSEIDON_PAIR_CTA=0
done
bash scripts/r2_sanitize.sh > /dev/null 2>&1
tail -n 4 gpurun_out/r2_sanitize_*.log >> $L
cat $L
