"""compute_image_mean: per-pixel mean of a record DB -> BlobProto file (mean.binaryproto).

    python -m poseidon_b200.tools.compute_image_mean INPUT_DB OUTPUT_FILE
reference: tools/compute_image_mean.cpp:17-150.
"""
from __future__ import annotations

import sys

import numpy as np

from .. import proto as P
from ..data.db import open_db


def compute_mean(db_path: str) -> np.ndarray:
    db = open_db(db_path)
    acc, n = None, 0
    for i in range(len(db)):
        d = db.datum(i)
        shape = (d.channels, d.height, d.width)
        if d.has("data") and len(d.data):
            x = np.frombuffer(d.data, dtype=np.uint8).astype(np.float64).reshape(shape)
        else:
            x = np.asarray(d.float_data, dtype=np.float64).reshape(shape)
        if acc is None:
            acc = np.zeros(shape, dtype=np.float64)
        elif acc.shape != shape:
            raise ValueError("Incorrect data field size")
        acc += x
        n += 1
    if n == 0:
        raise ValueError("empty database")
    return (acc / n).astype(np.float32)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 2:
        print(__doc__)
        return 1
    mean = compute_mean(argv[0])
    P.write_binary(argv[1], P.array_to_blob(mean[None]))
    print(f"Processed mean of shape {mean.shape} -> {argv[1]}", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
