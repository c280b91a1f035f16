"""Bösen-style table API (``PSTableGroup`` / ``Table``) on replicated device tensors with SSP clocks; row types
(dense / fp16-wire / sparse / sorted-vector-map / sparse-feature / multiplicative) and server table logic (AdaRevision)."""
from . import rows  # noqa: F401
from .rows import (AbstractRow, DenseRow, DenseRowFloat16, MultiplicativeDenseRow, SortedVectorMapRow, SparseFeatureRow,  # noqa: F401
                   SparseRow, register_row)
from .table import ConsistencyModel, PSTableGroup, RowTable, Table, VectorClock  # noqa: F401
from .table_logic import AbstractServerTableLogic, AdaRevisionServerTableLogic  # noqa: F401
