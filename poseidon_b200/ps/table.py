"""A serverless re-imagining of the Bösen table API.

The reference shards every table over server threads on all machines and moves rows through host oplogs and
ZeroMQ (ps/src/petuum_ps_common/include/ps_table_group.hpp:29-138, table.hpp:108-193,
ps/src/petuum_ps/client/table_group.cpp, consistency/ssp*_consistency_controller.cpp).  Here a table is a dense
device tensor replicated on every rank:

* ``inc / batch_inc / dense_batch_inc`` apply to the local replica at once (read-my-writes) and accumulate in an
  oplog tensor;
* ``PSTableGroup.clock()`` snapshots each table's oplog and starts an asynchronous all-reduce (NCCL on GPUs, gloo
  on CPU); when it completes, the *other* workers' part (sum − own) is folded into the replica — every update is
  applied exactly once everywhere;
* ``get(row, clock)`` enforces SSP: before serving a read at worker-clock ``c`` with staleness ``s`` all oplogs of
  clocks ``<= c - s - 1`` are folded in (``s = 0`` = BSP, the SSPPush setting of every shipped script);
* ``global_barrier()`` = ``s + 1`` clocks, as in the reference (table_group.cpp:200-204).
"""
from __future__ import annotations

import enum
from collections import deque
from typing import Dict, Optional

import torch
import torch.distributed as dist


class ConsistencyModel(enum.Enum):
    SSP = 0
    SSPPush = 1
    SSPAggr = 2
    LocalOOC = 3


class VectorClock:
    """min-clock over a fixed set of participants (ps/src/petuum_ps_common/util/vector_clock.cpp:19-66)."""

    def __init__(self, ids=()):
        self.clk: Dict[int, int] = {i: 0 for i in ids}
        self.min_clock = 0

    def add_clock(self, i: int, clock: int = 0):
        self.clk[i] = clock
        self.min_clock = min(self.clk.values())

    def tick(self, i: int) -> int:
        """Advance participant i; returns the new min clock if it moved, else 0."""
        self.clk[i] += 1
        new_min = min(self.clk.values())
        if new_min != self.min_clock:
            self.min_clock = new_min
            return new_min
        return 0

    def get_clock(self, i: int) -> int:
        return self.clk[i]

    def get_min_clock(self) -> int:
        return self.min_clock


class Table:
    def __init__(self, group: "PSTableGroup", table_id: int, num_rows: int, row_capacity: int, dtype, staleness: int):
        self.group, self.id = group, table_id
        self.staleness = staleness
        dev = group.device
        self.data = torch.zeros(num_rows, row_capacity, dtype=dtype, device=dev)
        self.oplog = torch.zeros_like(self.data)
        self.dirty = False
        self.inflight: deque = deque()          # (clock, total, own, work)

    # ---- writes -------------------------------------------------------------------------------------
    def inc(self, row_id: int, column_id: int, delta):
        self.data[row_id, column_id] += delta
        self.oplog[row_id, column_id] += delta
        self.dirty = True

    def batch_inc(self, row_id: int, updates):
        """updates: {col: delta} (UpdateBatch) or a dense 1-D tensor for the whole row (DenseUpdateBatch)."""
        if isinstance(updates, dict):
            cols = torch.tensor(list(updates.keys()), device=self.data.device)
            vals = torch.tensor(list(updates.values()), dtype=self.data.dtype, device=self.data.device)
            self.data[row_id].index_add_(0, cols, vals)
            self.oplog[row_id].index_add_(0, cols, vals)
        else:
            u = torch.as_tensor(updates, dtype=self.data.dtype, device=self.data.device)
            self.data[row_id, : u.numel()] += u
            self.oplog[row_id, : u.numel()] += u
        self.dirty = True

    # ---- reads --------------------------------------------------------------------------------------
    def get(self, row_id: int, clock: Optional[int] = None) -> torch.Tensor:
        c = self.group.clock_value if clock is None else clock
        self._fold(min_required=c - self.staleness - 1)
        return self.data[row_id]

    def dense_batch_inc(self, row_id: int, values, index_st: int = 0):
        """Table::DenseBatchInc: add a contiguous run of deltas starting at column ``index_st``."""
        values = torch.as_tensor(values)
        self.batch_inc(row_id, {index_st + i: float(v) for i, v in enumerate(values.reshape(-1).tolist())})

    def get_async(self, row_id: int):
        return self.get_async_forced(row_id)

    # thread-cache variants of the reference (ThreadGet / ThreadInc / FlushThreadCache): one cache level here
    def thread_get(self, row_id: int, clock: Optional[int] = None):
        return self.get(row_id, clock)

    def thread_inc(self, row_id: int, column_id: int, delta):
        return self.inc(row_id, column_id, delta)

    def flush_thread_cache(self):
        return None

    def get_async_forced(self, row_id: int):      # subscription is implicit (full replica)
        return None

    def wait_pending_async_get(self):
        return None

    # ---- clock machinery ----------------------------------------------------------------------------
    def _clock(self, clock: int):
        if self.dirty or self.group.world > 1:
            own = self.oplog.clone()
            total = own.clone()
            work = dist.all_reduce(total, async_op=True) if self.group.world > 1 else None
            self.inflight.append((clock, total, own, work))
            self.oplog.zero_()
            self.dirty = False

    def _fold(self, min_required: int):
        while self.inflight:
            clk, total, own, work = self.inflight[0]
            done = work is None or work.is_completed()
            if not done and clk > min_required:
                break
            if work is not None:
                work.wait()
            self.data += total - own
            self.inflight.popleft()


class PSTableGroup:
    """Process-level singleton facade with the reference's static API names."""
    _inst: Optional["PSTableGroup"] = None

    def __init__(self, rank_ctx=None, staleness: int = 0, consistency_model=ConsistencyModel.SSPPush):
        self.rank = rank_ctx.rank if rank_ctx is not None else 0
        self.world = rank_ctx.world_size if rank_ctx is not None else 1
        self.device = rank_ctx.device if rank_ctx is not None else torch.device("cpu")
        self.staleness = staleness
        self.consistency_model = consistency_model
        self.tables: Dict[int, Table] = {}
        self.clock_value = 0
        self.tables_created = False

    # -- reference-style static entry points ------------------------------------------------------------
    @classmethod
    def init(cls, rank_ctx=None, staleness: int = 0, **kw) -> "PSTableGroup":
        cls._inst = PSTableGroup(rank_ctx, staleness, **kw)
        return cls._inst

    @classmethod
    def instance(cls) -> "PSTableGroup":
        if cls._inst is None:
            raise RuntimeError("PSTableGroup.init() has not been called")
        return cls._inst

    def create_table(self, table_id: int, num_rows: int, row_capacity: int, dtype=torch.float32,
                     staleness: Optional[int] = None) -> Table:
        if self.tables_created:
            raise RuntimeError("CreateTableDone() was already called")
        if table_id in self.tables:
            raise ValueError(f"table {table_id} exists")
        t = Table(self, table_id, num_rows, row_capacity, dtype, self.staleness if staleness is None else staleness)
        self.tables[table_id] = t
        return t

    def create_table_done(self):
        self.tables_created = True
        self._barrier()

    def get_table_or_die(self, table_id: int) -> Table:
        if table_id not in self.tables:
            raise KeyError(f"table {table_id} does not exist")
        return self.tables[table_id]

    # -- API names that only matter with real server / background threads: accepted, nothing to do ------------------
    def register_row(self, row_type_id: int = 0, row_cls=None):
        """PSTableGroup::RegisterRow<...>: rows are plain tensors here."""
        return row_type_id

    def wait_thread_register(self):
        return None

    def turn_on_early_comm(self):
        """SSPPush/SSPAggr early communication toggle: updates are always exchanged at clock boundaries here."""
        self.early_comm = True

    def turn_off_early_comm(self):
        self.early_comm = False

    def register_thread(self):
        return 0

    def deregister_thread(self):
        return None

    def clock(self):
        """End of one worker iteration: flush oplogs, advance the clock."""
        for t in self.tables.values():
            t._clock(self.clock_value)
        self.clock_value += 1
        for t in self.tables.values():
            t._fold(min_required=self.clock_value - t.staleness - 1)

    def global_barrier(self):
        for _ in range(self.staleness + 1):
            self.clock()
        for t in self.tables.values():
            t._fold(min_required=self.clock_value)
        self._barrier()

    def _barrier(self):
        if self.world > 1:
            if self.device.type == "cuda":
                dist.barrier(device_ids=[self.device.index])
            else:
                dist.barrier()

    def shut_down(self):
        for t in self.tables.values():
            t._fold(min_required=self.clock_value)
        self._barrier()
        PSTableGroup._inst = None
