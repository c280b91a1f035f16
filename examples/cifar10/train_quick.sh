#!/bin/bash
# CIFAR-10 quick (reference: examples/cifar10/train_cifar10.sh).   examples/cifar10/train_quick.sh [NUM_GPUS]
set -e
cd "$(dirname "$0")/../.."
N=${1:-1}
python -m poseidon_b200.models.zoo --out models --only cifar10_quick
[ -d data/cifar10 ] && [ ! -d examples/cifar10/cifar10_train_db ] && \
  python -m poseidon_b200.tools.convert_cifar_data data/cifar10 examples/cifar10
python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29400 \
  -m poseidon_b200.tools.caffe_main train --solver=models/cifar10_quick/solver.prototxt --net_outputs=output/cifar10_quick "${@:2}"
