"""convert_cifar_data: CIFAR-10 binary batches -> train / test record DBs (3x32x32 uint8 + label).

    python -m poseidon_b200.tools.convert_cifar_data INPUT_FOLDER OUTPUT_FOLDER

Reads ``data_batch_{1..5}.bin`` and ``test_batch.bin`` (1 label byte + 3072 pixel bytes per record) and writes
``OUTPUT_FOLDER/cifar10_train_db`` and ``cifar10_test_db``.
reference: examples/cifar10/convert_cifar_data.cpp (kCIFARBatchSize 10000, key = %05d).
"""
from __future__ import annotations

import os
import sys

import numpy as np

REC = 1 + 3 * 32 * 32


def _read_batch(path):
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size % REC:
        raise ValueError(f"{path}: size {raw.size} is not a multiple of {REC}")
    raw = raw.reshape(-1, REC)
    return raw[:, 0], raw[:, 1:]


def convert(input_folder: str, output_folder: str):
    from .. import proto as P
    from ..data.db import RecordWriter
    counts = {}
    for name, files in (("cifar10_train_db", [f"data_batch_{i}.bin" for i in range(1, 6)]),
                        ("cifar10_test_db", ["test_batch.bin"])):
        n = 0
        with RecordWriter(os.path.join(output_folder, name, "data.pdb")) as w:
            for fn in files:
                path = os.path.join(input_folder, fn)
                if not os.path.exists(path):
                    raise FileNotFoundError(f"Unable to open train file {path}")
                labels, pix = _read_batch(path)
                for i in range(len(labels)):
                    d = P.Datum(channels=3, height=32, width=32, label=int(labels[i]))
                    d.data = pix[i].tobytes()
                    w.put(f"{n:05d}", d.SerializeToString())
                    n += 1
        counts[name] = n
    return counts


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__)
        return 1
    for k, v in convert(*argv).items():
        print(f"{k}: {v} records")
    return 0


if __name__ == "__main__":
    sys.exit(main())
