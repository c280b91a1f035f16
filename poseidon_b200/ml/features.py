"""Feature containers (reference: ps/src/ml/feature/{abstract,dense,sparse}_feature.hpp).

``DenseFeature`` wraps a 1-D float tensor, ``SparseFeature`` a sorted (ids, values) pair with binary-search lookup,
``SparseBatch`` a CSR block of samples (what the loaders return) that converts to ``torch.sparse_csr`` / dense."""
from __future__ import annotations

from typing import Iterator, Tuple

import torch


class DenseFeature:
    def __init__(self, values, dtype=torch.float32):
        self.v = torch.as_tensor(values, dtype=dtype).reshape(-1).clone()

    @property
    def feature_dim(self) -> int:
        return self.v.numel()

    num_entries = feature_dim

    def __getitem__(self, i: int) -> float:
        return float(self.v[i])

    def __setitem__(self, i: int, x: float):
        self.v[i] = x

    def entries(self) -> Iterator[Tuple[int, float]]:
        return ((i, float(x)) for i, x in enumerate(self.v))

    def to_dense(self) -> torch.Tensor:
        return self.v

    def __repr__(self):
        return " ".join(f"{i}:{x:g}" for i, x in self.entries())


class SparseFeature:
    """Sorted ids + values; ``feature_dim`` bounds the ids (reference: sparse_feature.hpp:17-60)."""

    def __init__(self, ids, values, feature_dim: int):
        ids = torch.as_tensor(ids, dtype=torch.int64).reshape(-1)
        values = torch.as_tensor(values, dtype=torch.float32).reshape(-1)
        if ids.numel() != values.numel():
            raise ValueError("ids and values differ in length")
        order = torch.argsort(ids)
        self.ids, self.vals = ids[order].clone(), values[order].clone()
        self.feature_dim = int(feature_dim)
        if self.ids.numel() and (int(self.ids[0]) < 0 or int(self.ids[-1]) >= self.feature_dim):
            raise ValueError(f"feature id outside [0, {self.feature_dim})")
        if self.ids.numel() > 1 and bool((self.ids[1:] == self.ids[:-1]).any()):
            raise ValueError("duplicate feature id")

    @property
    def num_entries(self) -> int:
        return self.ids.numel()

    def _find(self, i: int) -> int:
        k = int(torch.searchsorted(self.ids, torch.tensor(i)))
        return k if k < self.ids.numel() and int(self.ids[k]) == i else -1

    def __getitem__(self, i: int) -> float:
        k = self._find(i)
        return float(self.vals[k]) if k >= 0 else 0.0

    def __setitem__(self, i: int, x: float):
        if not 0 <= i < self.feature_dim:
            raise IndexError(i)
        k = self._find(i)
        if k >= 0:
            self.vals[k] = x
            return
        pos = int(torch.searchsorted(self.ids, torch.tensor(i)))
        self.ids = torch.cat([self.ids[:pos], torch.tensor([i]), self.ids[pos:]])
        self.vals = torch.cat([self.vals[:pos], torch.tensor([float(x)]), self.vals[pos:]])

    def entries(self) -> Iterator[Tuple[int, float]]:
        return ((int(i), float(x)) for i, x in zip(self.ids, self.vals))

    def to_dense(self) -> torch.Tensor:
        out = torch.zeros(self.feature_dim)
        out[self.ids] = self.vals
        return out

    def __repr__(self):
        return " ".join(f"{i}:{x:g}" for i, x in self.entries())


class SparseBatch:
    """CSR block: row r holds ids ``indices[indptr[r]:indptr[r+1]]``."""

    def __init__(self, indptr, indices, values, feature_dim: int):
        self.indptr = torch.as_tensor(indptr, dtype=torch.int64)
        self.indices = torch.as_tensor(indices, dtype=torch.int64)
        self.values = torch.as_tensor(values, dtype=torch.float32)
        self.feature_dim = int(feature_dim)
        if self.indices.numel() and int(self.indices.max()) >= self.feature_dim:
            raise ValueError(f"feature id {int(self.indices.max())} >= feature_dim {self.feature_dim}")

    def __len__(self) -> int:
        return self.indptr.numel() - 1

    def __getitem__(self, r: int) -> SparseFeature:
        a, b = int(self.indptr[r]), int(self.indptr[r + 1])
        return SparseFeature(self.indices[a:b], self.values[a:b], self.feature_dim)

    def to_sparse_csr(self) -> torch.Tensor:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)       # "sparse CSR is beta" / "invariant checks implicit"
            return torch.sparse_csr_tensor(self.indptr, self.indices, self.values,
                                           size=(len(self), self.feature_dim), check_invariants=True)

    def to_dense(self) -> torch.Tensor:
        out = torch.zeros(len(self), self.feature_dim)
        rows = torch.repeat_interleave(torch.arange(len(self)), self.indptr[1:] - self.indptr[:-1])
        out[rows, self.indices] = self.values
        return out

    def matvec(self, w: torch.Tensor) -> torch.Tensor:
        """X·w for a dense weight vector (the inner loop of the linear models built on this library)."""
        prod = self.values * w[self.indices]
        rows = torch.repeat_interleave(torch.arange(len(self)), self.indptr[1:] - self.indptr[:-1])
        return torch.zeros(len(self), dtype=prod.dtype).index_add_(0, rows, prod)
