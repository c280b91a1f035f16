"""Row types of the table API — the reference's ``petuum_ps_common/storage`` family on tensors.

The reference keeps every row as a C++ object behind ``AbstractRow`` (ApplyInc / ApplyBatchInc / ApplyDenseBatchInc,
thread-safe and *Unsafe variants, CopyToVector): ``dense_row.hpp`` (vector store), ``dense_row_float16.hpp`` (fp32 math,
half on the wire), ``sparse_row.hpp:12`` (``std::map`` store), ``sorted_vector_map_row.hpp`` (sorted (col, value) vector
that drops entries reaching zero), ``sparse_feature_row.hpp`` (sorted vector, bulk read for ML features),
``multiplicative_dense_row.hpp`` (Inc multiplies).  Here a row is a small tensor-backed object with the same operations;
sparse rows keep a sorted int64 column vector + a value vector (binary search reads, merge-coalesce writes), so whole
oplogs are applied with vectorised tensor ops and the same code runs on the CPU and on a GPU.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple, Type

import torch


class AbstractRow:
    """reference: ps/src/petuum_ps_common/include/abstract_row.hpp."""
    sparse = False
    wire_dtype = None            # dtype used when updates of this row type cross the wire (None = the table's dtype)

    def __init__(self, capacity: int, dtype=torch.float32, device="cpu"):
        self.capacity, self.dtype, self.device = int(capacity), dtype, torch.device(device)

    # -- writes ---------------------------------------------------------------------------------------
    def apply_inc(self, col: int, delta) -> None:
        self.apply_batch_inc([int(col)], [delta])

    def apply_batch_inc(self, cols, deltas) -> None:
        raise NotImplementedError

    def apply_dense_batch_inc(self, deltas, index_st: int = 0) -> None:
        d = torch.as_tensor(deltas, dtype=self.dtype, device=self.device).reshape(-1)
        self.apply_batch_inc(torch.arange(index_st, index_st + d.numel(), device=self.device), d)

    # -- reads ----------------------------------------------------------------------------------------
    def __getitem__(self, col: int):
        raise NotImplementedError

    def copy_to_vector(self) -> List[Tuple[int, float]]:
        raise NotImplementedError

    def to_dense(self) -> torch.Tensor:
        raise NotImplementedError

    def num_entries(self) -> int:
        raise NotImplementedError

    def _cols_vals(self, cols, deltas):
        c = torch.as_tensor(cols, dtype=torch.int64, device=self.device).reshape(-1)
        v = torch.as_tensor(deltas, dtype=self.dtype, device=self.device).reshape(-1)
        if c.numel() and (int(c.min()) < 0 or int(c.max()) >= self.capacity):
            raise IndexError(f"column id out of range [0, {self.capacity})")
        return c, v


class DenseRow(AbstractRow):
    """reference: storage/dense_row.hpp (VectorStore)."""

    def __init__(self, capacity, dtype=torch.float32, device="cpu"):
        super().__init__(capacity, dtype, device)
        self.data = torch.zeros(self.capacity, dtype=dtype, device=self.device)

    def apply_batch_inc(self, cols, deltas):
        c, v = self._cols_vals(cols, deltas)
        self.data.index_add_(0, c, v)

    def __getitem__(self, col):
        return self.data[col]

    def copy_to_vector(self):
        return list(enumerate(self.data.tolist()))

    def to_dense(self):
        return self.data

    def num_entries(self):
        return self.capacity


class DenseRowFloat16(DenseRow):
    """fp32 at rest, 16-bit on the wire (storage/dense_row_float16.hpp, vector_store_float16.hpp)."""
    wire_dtype = torch.float16


class MultiplicativeDenseRow(DenseRow):
    """Inc multiplies instead of adding (storage/multiplicative_dense_row.hpp); rows start at 1."""

    def __init__(self, capacity, dtype=torch.float32, device="cpu"):
        super().__init__(capacity, dtype, device)
        self.data.fill_(1)

    def apply_batch_inc(self, cols, deltas):
        c, v = self._cols_vals(cols, deltas)
        # duplicates multiply too: fold them first (log-free: sequential products per unique column)
        uniq, inv = torch.unique(c, return_inverse=True)
        prod = torch.ones(uniq.numel(), dtype=self.dtype, device=self.device)
        prod.scatter_reduce_(0, inv, v, reduce="prod")
        self.data[uniq] *= prod


class SparseRow(AbstractRow):
    """col -> value map (storage/sparse_row.hpp:12, map_store.hpp): unset columns read as 0, entries persist at 0."""
    sparse = True
    drop_zeros = False

    def __init__(self, capacity, dtype=torch.float32, device="cpu"):
        super().__init__(capacity, dtype, device)
        self.cols = torch.empty(0, dtype=torch.int64, device=self.device)
        self.vals = torch.empty(0, dtype=dtype, device=self.device)

    def apply_batch_inc(self, cols, deltas):
        c, v = self._cols_vals(cols, deltas)
        if not c.numel():
            return
        allc = torch.cat([self.cols, c])
        allv = torch.cat([self.vals, v])
        uniq, inv = torch.unique(allc, return_inverse=True)            # sorted
        vals = torch.zeros(uniq.numel(), dtype=self.dtype, device=self.device)
        vals.index_add_(0, inv, allv)
        if self.drop_zeros:
            keep = vals != 0
            uniq, vals = uniq[keep], vals[keep]
        self.cols, self.vals = uniq, vals

    def __getitem__(self, col):
        col = int(col)
        i = int(torch.searchsorted(self.cols, torch.tensor([col], device=self.device))[0]) if self.cols.numel() else 0
        if i < self.cols.numel() and int(self.cols[i]) == col:
            return self.vals[i]
        return torch.zeros((), dtype=self.dtype, device=self.device)

    def copy_to_vector(self):
        return list(zip(self.cols.tolist(), self.vals.tolist()))

    def to_dense(self):
        d = torch.zeros(self.capacity, dtype=self.dtype, device=self.device)
        d[self.cols] = self.vals
        return d

    def num_entries(self):
        return int(self.cols.numel())


class SortedVectorMapRow(SparseRow):
    """Sorted (col, value) vector whose entries disappear when they reach zero — the LDA-style count row
    (storage/sorted_vector_map_row.hpp, sorted_vector_map_store.hpp)."""
    drop_zeros = True


class SparseFeatureRow(SparseRow):
    """Sorted-vector row for sparse float features with bulk reads (storage/sparse_feature_row.hpp, sorted_vector_store.hpp)."""

    def copy_to_tensors(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.cols.clone(), self.vals.clone()


ROW_TYPES: Dict[int, Type[AbstractRow]] = {}
DENSE_FLOAT_ROW, SPARSE_ROW, SORTED_VECTOR_MAP_ROW, SPARSE_FEATURE_ROW, DENSE_FLOAT16_ROW, MULTIPLICATIVE_DENSE_ROW = range(6)


def register_row(row_type_id: int, row_cls: Type[AbstractRow]) -> int:
    """PSTableGroup::RegisterRow<ROW>(row_type_id) (ps_table_group.hpp:67-76): the id is what CreateTable refers to."""
    if not (isinstance(row_cls, type) and issubclass(row_cls, AbstractRow)):
        raise TypeError("row classes derive from AbstractRow")
    if row_type_id in ROW_TYPES and ROW_TYPES[row_type_id] is not row_cls:
        raise ValueError(f"row type id {row_type_id} is already registered as {ROW_TYPES[row_type_id].__name__}")
    ROW_TYPES[row_type_id] = row_cls
    return row_type_id


for _i, _c in ((DENSE_FLOAT_ROW, DenseRow), (SPARSE_ROW, SparseRow), (SORTED_VECTOR_MAP_ROW, SortedVectorMapRow),
               (SPARSE_FEATURE_ROW, SparseFeatureRow), (DENSE_FLOAT16_ROW, DenseRowFloat16),
               (MULTIPLICATIVE_DENSE_ROW, MultiplicativeDenseRow)):
    register_row(_i, _c)


def coalesce(rows: Iterable[int], cols: Iterable[int], vals: torch.Tensor, capacity: int):
    """Sum duplicate (row, col) entries of a COO update list; returns sorted (rows, cols, vals) tensors."""
    r = torch.as_tensor(rows, dtype=torch.int64).reshape(-1)
    c = torch.as_tensor(cols, dtype=torch.int64).reshape(-1)
    v = torch.as_tensor(vals).reshape(-1)
    if not r.numel():
        return r, c, v
    key = r * capacity + c
    uniq, inv = torch.unique(key, return_inverse=True)
    out = torch.zeros(uniq.numel(), dtype=v.dtype, device=v.device)
    out.index_add_(0, inv.to(v.device), v)
    return uniq // capacity, uniq % capacity, out
