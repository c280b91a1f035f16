#!/bin/bash
# GoogLeNet quick solver (reference: examples/googlenet/train_googlenet.sh, num_rows_per_table=32, svb=true).
#   examples/googlenet/train_googlenet.sh [NUM_GPUS] [extra flags]
set -e
cd "$(dirname "$0")/../.."
N=${1:-8}
python -m poseidon_b200.models.zoo --out models --only googlenet
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29402 \
  -m poseidon_b200.tools.caffe_main train --solver=models/googlenet/solver.prototxt --svb=true \
  --net_outputs=output/googlenet "${@:2}"
