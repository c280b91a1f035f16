#!/bin/bash
# LeNet on MNIST (reference: examples/mnist/train_mnist.sh + lenet_solver.prototxt).
#   examples/mnist/train_lenet.sh [cpu|gpu]
set -e
cd "$(dirname "$0")/../.."
MODE=${1:-cpu}
python -m poseidon_b200.models.zoo --out models --only lenet
if [ -d data/mnist ] && [ ! -d examples/mnist/mnist_train_lmdb ]; then
  python -m poseidon_b200.tools.convert_mnist_data data/mnist/train-images-idx3-ubyte data/mnist/train-labels-idx1-ubyte examples/mnist/mnist_train_lmdb
  python -m poseidon_b200.tools.convert_mnist_data data/mnist/t10k-images-idx3-ubyte data/mnist/t10k-labels-idx1-ubyte examples/mnist/mnist_test_lmdb
fi
GPU=""; [ "$MODE" == "gpu" ] && GPU="--gpu=0"
python -m poseidon_b200.tools.caffe_main train --solver=models/lenet/solver.prototxt $GPU \
  --net_outputs=output/lenet "${@:2}"
