#!/bin/bash
# MNIST idx files (data/mnist/*-ubyte) -> train / test LMDBs (reference: examples/mnist/create_mnist.sh).
set -e
cd "$(dirname "$0")/../.."
DATA=${1:-data/mnist}
OUT=examples/mnist
python -m poseidon_b200.tools.convert_mnist_data "$DATA/train-images-idx3-ubyte" "$DATA/train-labels-idx1-ubyte" "$OUT/mnist_train_lmdb"
python -m poseidon_b200.tools.convert_mnist_data "$DATA/t10k-images-idx3-ubyte" "$DATA/t10k-labels-idx1-ubyte" "$OUT/mnist_test_lmdb"
echo "Done: $OUT/mnist_{train,test}_lmdb"
