#!/bin/bash
# Distributed feature extraction (reference: examples/feature_extraction/readme.md, tools/extract_features.cpp).
#   examples/feature_extraction/extract.sh WEIGHTS.caffemodel NET.prototxt fc7,prob out_fc7,out_prob NUM_BATCHES [NUM_GPUS]
set -e
cd "$(dirname "$0")/../.."
N=${6:-1}
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29403 \
  -m poseidon_b200.tools.extract_features "$1" "$2" "$3" "$4" "$5"
