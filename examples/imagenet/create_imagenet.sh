#!/bin/bash
# ILSVRC12 image folders + label lists -> train / val LMDBs, images resized to 256x256
# (reference: examples/imagenet/create_imagenet.sh).  Set the four paths below.
set -e
cd "$(dirname "$0")/../.."
TRAIN_ROOT=${TRAIN_ROOT:-/path/to/imagenet/train/}
VAL_ROOT=${VAL_ROOT:-/path/to/imagenet/val/}
LISTS=${LISTS:-data/ilsvrc12}              # train.txt / val.txt: "<relative path> <label>" per line
OUT=${OUT:-.}
for d in "$TRAIN_ROOT" "$VAL_ROOT"; do
  [ -d "$d" ] || { echo "Error: $d is not a directory (set TRAIN_ROOT / VAL_ROOT)"; exit 1; }
done
python -m poseidon_b200.tools.convert_imageset "$TRAIN_ROOT" "$LISTS/train.txt" "$OUT/ilsvrc12_train_lmdb" \
  --resize_height 256 --resize_width 256 --shuffle --backend lmdb
python -m poseidon_b200.tools.convert_imageset "$VAL_ROOT" "$LISTS/val.txt" "$OUT/ilsvrc12_val_lmdb" \
  --resize_height 256 --resize_width 256 --shuffle --backend lmdb
echo "Done."
