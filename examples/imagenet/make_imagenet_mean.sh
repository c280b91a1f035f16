#!/bin/bash
# Mean image of the training LMDB (reference: examples/imagenet/make_imagenet_mean.sh).
set -e
cd "$(dirname "$0")/../.."
mkdir -p data/ilsvrc12
python -m poseidon_b200.tools.compute_image_mean "${1:-ilsvrc12_train_lmdb}" data/ilsvrc12/imagenet_mean.binaryproto
echo "Done: data/ilsvrc12/imagenet_mean.binaryproto"
