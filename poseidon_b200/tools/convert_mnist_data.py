"""convert_mnist_data: MNIST idx files -> record DB of ``Datum`` (1x28x28 uint8 + label).

    python -m poseidon_b200.tools.convert_mnist_data train-images-idx3-ubyte train-labels-idx1-ubyte out_db

reference: examples/mnist/convert_mnist_data.cpp (magic 2051 / 2049 headers, big-endian counts, key = %08d).
"""
from __future__ import annotations

import gzip
import struct
import sys

import numpy as np


def _open(path):
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def read_idx(image_path: str, label_path: str):
    with _open(image_path) as f:
        magic, n, rows, cols = struct.unpack(">IIII", f.read(16))
        if magic != 2051:
            raise ValueError(f"{image_path}: incorrect image file magic {magic}")
        images = np.frombuffer(f.read(n * rows * cols), dtype=np.uint8).reshape(n, rows, cols)
    with _open(label_path) as f:
        magic, nl = struct.unpack(">II", f.read(8))
        if magic != 2049:
            raise ValueError(f"{label_path}: incorrect label file magic {magic}")
        labels = np.frombuffer(f.read(nl), dtype=np.uint8)
    if n != nl:
        raise ValueError("image / label count mismatch")
    return images, labels


def convert(image_path: str, label_path: str, db_path: str) -> int:
    from .. import proto as P
    from ..data.db import RecordWriter
    images, labels = read_idx(image_path, label_path)
    with RecordWriter(db_path if db_path.endswith(".pdb") else db_path + "/data.pdb") as w:
        for i in range(len(images)):
            d = P.Datum(channels=1, height=images.shape[1], width=images.shape[2], label=int(labels[i]))
            d.data = images[i].tobytes()
            w.put(f"{i:08d}", d.SerializeToString())
    return len(images)


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 3:
        print(__doc__)
        return 1
    n = convert(*argv)
    print(f"Processed {n} files.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
