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

from . import rows as R
from .table_logic import AbstractServerTableLogic


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
    """Dense float tables keep the fast path (one tensor, oplog tensor, asynchronous all-reduce).  Tables created with a
    sparse row type (``rows.SPARSE_ROW``, ``SORTED_VECTOR_MAP_ROW``, ``SPARSE_FEATURE_ROW``), a non-additive row type or a
    server table logic (AdaRevision) keep one row object per row id and exchange their oplogs as (row, col, value)
    triplets at the clock boundary — see :class:`RowTable`."""

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

    # Asynchronous reads (Table::GetAsync / GetAsyncForced / WaitPendingAsyncGet): the row is replicated, so "fetching"
    # means folding in every completed remote update without blocking; Wait blocks for the SSP-required ones.
    def get_async(self, row_id: int):
        self._fold(min_required=-(1 << 60))
        self._pending_async = True

    def get_async_forced(self, row_id: int):
        return self.get_async(row_id)

    def wait_pending_async_get(self):
        if getattr(self, "_pending_async", False):
            self._fold(min_required=self.group.clock_value - self.staleness - 1)
            self._pending_async = False

    # Thread-level cache (ThreadGet / ThreadInc / FlushThreadCache, client_table.cpp: thread_cache_): increments are
    # buffered per calling thread and become visible to the process (and the oplog) on flush or at the clock.
    def thread_get(self, row_id: int, clock: Optional[int] = None):
        row = self.get(row_id, clock).clone()
        for (r, c), d in self._tcache().items():
            if r == row_id:
                row[c] += d
        return row

    def thread_inc(self, row_id: int, column_id: int, delta):
        tc = self._tcache()
        tc[(row_id, column_id)] = tc.get((row_id, column_id), 0.0) + float(delta)

    def flush_thread_cache(self):
        tc = self._tcache()
        for (r, c), d in tc.items():
            self.inc(r, c, d)
        tc.clear()

    def _tcache(self) -> dict:
        import threading
        caches = self.__dict__.setdefault("_thread_caches", {})
        return caches.setdefault(threading.get_ident(), {})

    # ---- clock machinery ----------------------------------------------------------------------------
    def _clock(self, clock: int):
        for tc in self.__dict__.get("_thread_caches", {}).values():       # a clock flushes every thread cache
            for (r, c), d in tc.items():
                self.inc(r, c, d)
            tc.clear()
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


class RowTable:
    """Table of row OBJECTS (ps/rows.py): sparse / sorted-vector / multiplicative / fp16-wire rows, optionally with a
    server table logic.  Oplogs are (row, col, value) triplets; at every clock the triplets of all workers are exchanged
    (variable-length all-gather) and applied in worker order, so every replica ends each clock with identical rows.

    Without a table logic, own writes are applied locally at once (read-my-writes) and only the OTHER workers' triplets
    are applied at the fold; with a logic (AdaRevision) every batch, own included, goes through the logic at the fold.
    reference: client_table.cpp:26-158 (oplog partitioning), oplog/*.hpp (sparse / dense row oplogs), server_table.cpp."""

    def __init__(self, group: "PSTableGroup", table_id: int, num_rows: int, row_capacity: int, dtype, staleness: int,
                 row_type: int, table_logic: Optional[AbstractServerTableLogic] = None):
        self.group, self.id, self.staleness = group, table_id, staleness
        self.num_rows, self.row_capacity, self.dtype = num_rows, row_capacity, dtype
        cls = R.ROW_TYPES[row_type]
        self.row_cls = cls
        self.rows = [cls(row_capacity, dtype, group.device) for _ in range(num_rows)]
        self.logic = table_logic
        if self.logic is not None:
            self.logic.init(self)
            for i, r in enumerate(self.rows):
                self.logic.server_row_created(i, r)
        self.op_rows: list = []
        self.op_cols: list = []
        self.op_vals: list = []
        self.pending: deque = deque()           # (clock, [per-worker (rows, cols, vals, version)])
        self.version = 0                        # number of clocks folded in = version of the rows a reader sees
        self.read_version = 0

    # ---- writes -------------------------------------------------------------------------------------
    def _log(self, row_id, cols, vals):
        self.op_rows.append(torch.full((len(cols),), int(row_id), dtype=torch.int64))
        self.op_cols.append(torch.as_tensor(cols, dtype=torch.int64).reshape(-1).cpu())
        self.op_vals.append(torch.as_tensor(vals, dtype=torch.float32).reshape(-1).cpu())

    def inc(self, row_id: int, column_id: int, delta):
        self.batch_inc(row_id, {int(column_id): delta})

    def batch_inc(self, row_id: int, updates):
        if isinstance(updates, dict):
            cols, vals = list(updates.keys()), [float(v) for v in updates.values()]
        else:
            u = torch.as_tensor(updates).reshape(-1)
            cols, vals = list(range(u.numel())), u.tolist()
        if self.logic is None:
            self.rows[row_id].apply_batch_inc(cols, vals)                # read-my-writes
        self._log(row_id, cols, vals)

    def dense_batch_inc(self, row_id: int, values, index_st: int = 0):
        v = torch.as_tensor(values).reshape(-1)
        self.batch_inc(row_id, {index_st + i: float(x) for i, x in enumerate(v.tolist())})

    # ---- reads --------------------------------------------------------------------------------------
    def get(self, row_id: int, clock: Optional[int] = None):
        c = self.group.clock_value if clock is None else clock
        self._fold(min_required=c - self.staleness - 1)
        self.read_version = self.version
        return self.rows[row_id]

    # ---- clock machinery ----------------------------------------------------------------------------
    def _clock(self, clock: int):
        if self.op_rows:
            r, c, v = R.coalesce(torch.cat(self.op_rows), torch.cat(self.op_cols), torch.cat(self.op_vals), self.row_capacity)
        else:
            r = c = torch.empty(0, dtype=torch.int64)
            v = torch.empty(0, dtype=torch.float32)
        self.op_rows, self.op_cols, self.op_vals = [], [], []
        wire = self.row_cls.wire_dtype
        if wire is not None:
            vw = v.to(wire).float()                                      # what the peers will see ...
            if self.logic is None and r.numel():
                for rid in torch.unique(r).tolist():                     # ... and, from now on, this replica too
                    sel = r == rid
                    self.rows[rid].apply_batch_inc(c[sel], vw[sel] - v[sel])
            v = vw
        batches = self._exchange(r, c, v, self.read_version)
        self.pending.append((clock, batches))

    def _exchange(self, r, c, v, version):
        g = self.group
        if g.world == 1:
            return [(r, c, v, version)]
        dev = g.device if g.device.type == "cuda" else torch.device("cpu")
        n = torch.tensor([r.numel(), version], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n) for _ in range(g.world)]
        dist.all_gather(sizes, n)
        m = max(int(s[0]) for s in sizes)
        pad = torch.zeros(max(m, 1), 3, dtype=torch.float64, device=dev)
        if r.numel():
            pad[: r.numel(), 0], pad[: r.numel(), 1], pad[: r.numel(), 2] = r.to(dev), c.to(dev), v.to(dev)
        allp = [torch.zeros_like(pad) for _ in range(g.world)]
        dist.all_gather(allp, pad)
        out = []
        for q in range(g.world):
            k = int(sizes[q][0])
            t = allp[q][:k].cpu()
            out.append((t[:, 0].long(), t[:, 1].long(), t[:, 2].float(), int(sizes[q][1])))
        return out

    def _fold(self, min_required: int):
        """The exchange at the clock is synchronous, so every pending batch is already local: fold them all (SSP permits
        fresher-than-required reads).  Worker order inside a clock is fixed => identical rows on every replica."""
        while self.pending:
            _, batches = self.pending.popleft()
            for q, (r, c, v, ver) in enumerate(batches):
                if self.logic is None and q == self.group.rank:
                    continue                                             # own writes were applied at write time
                for rid in torch.unique(r).tolist():
                    sel = r == rid
                    if self.logic is None:
                        self.rows[rid].apply_batch_inc(c[sel], v[sel])
                    else:
                        last = not any((b[0] == rid).any() for b in batches[q + 1:])
                        self.logic.apply_row_oplog(rid, c[sel], v[sel], self.rows[rid], ver, end_of_version=last)
            self.version += 1
            if self.logic is not None:
                for rid in range(self.num_rows):
                    self.logic.server_row_sent(rid, self.version, self.group.world)


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
        self.app_threads: Dict[int, int] = {}
        self._next_thread_id = 0
        self.num_app_threads = 0
        self.vector_clock = VectorClock()
        self.early_comm = False

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
                     staleness: Optional[int] = None, row_type: int = R.DENSE_FLOAT_ROW,
                     table_logic: Optional[AbstractServerTableLogic] = None):
        """ClientTableConfig equivalent: ``row_type`` is an id registered with :meth:`register_row` (dense float rows take
        the tensor fast path), ``table_logic`` a server-side logic such as ``AdaRevisionServerTableLogic``."""
        if self.tables_created:
            raise RuntimeError("CreateTableDone() was already called")
        if table_id in self.tables:
            raise ValueError(f"table {table_id} exists")
        if row_type not in R.ROW_TYPES:
            raise KeyError(f"row type {row_type} was never registered (PSTableGroup.register_row)")
        st = self.staleness if staleness is None else staleness
        if row_type == R.DENSE_FLOAT_ROW and table_logic is None:
            t = Table(self, table_id, num_rows, row_capacity, dtype, st)
        else:
            t = RowTable(self, table_id, num_rows, row_capacity, dtype, st, row_type, table_logic)
        self.tables[table_id] = t
        return t

    def create_table_done(self):
        self.tables_created = True
        self._barrier()

    def get_table_or_die(self, table_id: int) -> Table:
        if table_id not in self.tables:
            raise KeyError(f"table {table_id} does not exist")
        return self.tables[table_id]

    # -- row types, application threads, early communication ----------------------------------------------------------
    def register_row(self, row_type_id: int = 0, row_cls=None):
        """PSTableGroup::RegisterRow<ROW>(id): makes a row class available to create_table(row_type=id)."""
        if row_cls is None:
            if row_type_id not in R.ROW_TYPES:
                raise KeyError(f"row type {row_type_id}: no class given and none registered")
            return row_type_id
        return R.register_row(row_type_id, row_cls)

    def register_thread(self) -> int:
        """RegisterThread: application threads get consecutive ids; the vector clock tracks them (table_group.cpp:160-176)."""
        import threading
        tid = threading.get_ident()
        if tid not in self.app_threads:
            self.app_threads[tid] = self._next_thread_id
            self._next_thread_id += 1
            self.vector_clock.add_clock(self.app_threads[tid], self.clock_value)
        return self.app_threads[tid]

    def deregister_thread(self):
        import threading
        tid = threading.get_ident()
        if tid in self.app_threads:
            for t in self.tables.values():
                if hasattr(t, "flush_thread_cache"):
                    t.flush_thread_cache()
            del self.app_threads[tid]

    def wait_thread_register(self):
        """Returns once `num_app_threads` threads have registered (table_group.cpp:150-158); immediate when none are expected."""
        import time
        deadline = time.time() + 60
        while len(self.app_threads) < self.num_app_threads:
            if time.time() > deadline:
                raise TimeoutError("wait_thread_register: application threads did not register")
            time.sleep(0.001)

    def turn_on_early_comm(self):
        """SSPPush / SSPAggr early communication: completed remote updates are folded in as soon as any table is touched
        (not only when the staleness bound forces it)."""
        self.early_comm = True

    def turn_off_early_comm(self):
        self.early_comm = False

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
