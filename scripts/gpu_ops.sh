#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/ops_tests.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $L 2>&1
for t in test_lrn test_pool test_softmax test_dropout test_colsum test_transform test_conv_fwd_bwd test_first_layer test_inner_product; do
  echo "=== $t" >> $L
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -k $t -p no:cacheprovider 2>&1 | tail -40 >> $L
done

echo "=== engine" >> $L
timeout 900 python -m pytest tests/test_engine_gpu.py -q -p no:cacheprovider -x 2>&1 | tail -60 >> $L
tail -80 $L
