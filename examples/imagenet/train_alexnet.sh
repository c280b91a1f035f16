#!/bin/bash
# AlexNet / CaffeNet on ILSVRC12 (reference: examples/imagenet/train_imagenet.sh — SSPPush, staleness 0, svb optional).
#   examples/imagenet/train_alexnet.sh [NUM_GPUS] [alexnet|caffenet] [extra caffe_main flags, e.g. --table_staleness=1]
# Expects ilsvrc12_train_lmdb / ilsvrc12_val_lmdb record DBs (tools.convert_imageset) and the mean file; without them the
# data layers generate synthetic 3x256x256 images, which is enough to reproduce the throughput numbers.
set -e
cd "$(dirname "$0")/../.."
N=${1:-8}
MODEL=${2:-alexnet}
python -m poseidon_b200.models.zoo --out models --only "$MODEL"
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29401 \
  -m poseidon_b200.tools.caffe_main train --solver=models/$MODEL/solver.prototxt --svb=true --table_staleness=0 \
  --net_outputs=output/$MODEL --stats_path=output/${MODEL}_stats "${@:3}"
