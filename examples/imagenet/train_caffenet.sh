#!/bin/bash
# CaffeNet on ILSVRC12 (reference: examples/imagenet/train_imagenet.sh with models/bvlc_reference_caffenet/solver.prototxt).
#   examples/imagenet/train_caffenet.sh [NUM_GPUS] [extra caffe_main flags, e.g. --svb=true --snapshot=...]
set -e
cd "$(dirname "$0")/../.."
N=${1:-8}
python -m poseidon_b200.models.zoo --out models --only caffenet
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29400 \
  -m poseidon_b200.tools.caffe_main train --solver=models/caffenet/solver.prototxt --svb=true \
  --net_outputs=output/caffenet "${@:2}"
