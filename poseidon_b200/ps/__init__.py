"""Bösen-style table API (``PSTableGroup`` / ``Table``) on replicated device tensors with SSP clocks."""
from .table import ConsistencyModel, PSTableGroup, Table, VectorClock  # noqa: F401
