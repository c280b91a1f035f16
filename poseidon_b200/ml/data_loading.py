"""Dataset readers of the ML helper library (reference: ps/src/ml/util/data_loading.{hpp,cpp}).

* dense binary: per sample ``int32 label`` followed by ``feature_dim`` float32 values;
* LibSVM text: ``label id:value id:value ...`` (optionally snappy-compressed as a whole), parsed by the C++ host module
  (csrc_host/libsvm_parser.cpp) on one thread per slice of the file;
* sparse binary: per sample ``int32 nnz | int32 label | nnz x int32 ids | nnz x float32 values``.

Loaders return ``(features, labels)`` with features a 2-D float tensor (dense) or a :class:`SparseBatch` (CSR)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .features import SparseBatch


def _maybe_snappy(raw: bytes, snappy_compressed: bool) -> bytes:
    if not snappy_compressed:
        return raw
    from ..data import native
    return native.module().snappy_uncompress(raw)


def read_data_label_binary(filename: str, feature_dim: int, num_data: int, feature_one_based: bool = False,
                           label_one_based: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(features [num_data, feature_dim] float32, labels [num_data] int32).  ``feature_one_based`` has no meaning for
    dense rows (kept for signature parity)."""
    rec = np.dtype([("label", "<i4"), ("x", "<f4", (feature_dim,))])
    arr = np.fromfile(filename, dtype=rec, count=num_data)
    if arr.shape[0] < num_data:
        raise IOError(f"{filename}: {arr.shape[0]} samples, {num_data} requested")
    labels = arr["label"].astype(np.int32) - (1 if label_one_based else 0)
    return torch.from_numpy(np.ascontiguousarray(arr["x"])), torch.from_numpy(labels)


def read_data_label_libsvm(filename: str, feature_dim: int, num_data: int = -1, feature_one_based: bool = False,
                           label_one_based: bool = False, snappy_compressed: bool = False, dense: bool = False,
                           threads: int = 0):
    """(SparseBatch | dense tensor, labels int32).  Reads at most ``num_data`` samples (all if negative)."""
    from ..data import native
    with open(filename, "rb") as f:
        raw = _maybe_snappy(f.read(), snappy_compressed)
    labels, indptr, indices, values = native.module().parse_libsvm(raw, feature_one_based, label_one_based,
                                                                   int(num_data), int(threads))
    batch = SparseBatch(indptr, indices.astype(np.int64), values, feature_dim)
    return (batch.to_dense() if dense else batch), torch.from_numpy(labels)


def read_data_label_sparse_feature_binary(filename: str, feature_dim: int, num_data: int = -1,
                                          feature_one_based: bool = False, label_one_based: bool = False,
                                          snappy_compressed: bool = False):
    with open(filename, "rb") as f:
        raw = _maybe_snappy(f.read(), snappy_compressed)
    words = np.frombuffer(raw, dtype="<i4")
    labels, indptr, ids, vals = [], [0], [], []
    p = 0
    while p < words.size and (num_data < 0 or len(labels) < num_data):
        nnz, lab = int(words[p]), int(words[p + 1])
        if nnz < 0 or p + 2 + 2 * nnz > words.size:
            raise IOError(f"{filename}: truncated sample at word {p}")
        labels.append(lab - (1 if label_one_based else 0))
        ids.append(words[p + 2: p + 2 + nnz].astype(np.int64) - (1 if feature_one_based else 0))
        vals.append(words[p + 2 + nnz: p + 2 + 2 * nnz].view("<f4"))
        indptr.append(indptr[-1] + nnz)
        p += 2 + 2 * nnz
    cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt))
    return (SparseBatch(np.array(indptr), cat(ids, np.int64), cat(vals, np.float32), feature_dim),
            torch.tensor(labels, dtype=torch.int32))


def write_sparse_feature_binary(filename: str, batch: SparseBatch, labels) -> None:
    """Inverse of :func:`read_data_label_sparse_feature_binary` (zero-based ids and labels)."""
    with open(filename, "wb") as f:
        for r in range(len(batch)):
            a, b = int(batch.indptr[r]), int(batch.indptr[r + 1])
            f.write(np.array([b - a, int(labels[r])], "<i4").tobytes())
            f.write(batch.indices[a:b].numpy().astype("<i4").tobytes())
            f.write(batch.values[a:b].numpy().astype("<f4").tobytes())
