#!/bin/bash
# CIFAR-10 binary batches (data/cifar10/data_batch_*.bin, test_batch.bin) -> train / test databases + mean image
# (reference: examples/cifar10/create_cifar10.sh: convert_cifar_data.bin + compute_image_mean).
set -e
cd "$(dirname "$0")/../.."
DATA=${1:-data/cifar10}
OUT=examples/cifar10
python -m poseidon_b200.tools.convert_cifar_data "$DATA" "$OUT"
python -m poseidon_b200.tools.compute_image_mean "$OUT/cifar10_train_leveldb" "$OUT/mean.binaryproto"
echo "Done: $OUT/cifar10_{train,test}_leveldb, $OUT/mean.binaryproto"
