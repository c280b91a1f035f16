"""Server-side table logic: what happens to an update batch when it reaches the authoritative copy of a row.

reference: ps/src/petuum_ps_common/include/abstract_server_table_logic.hpp (hook), ps/src/petuum_ps/server/server_table.cpp:83-94
(where the server calls it instead of a plain BatchInc), ps/src/petuum_ps/server/adarevision_server_table_logic.hpp:38-75 +
.cpp:52-170 (AdaRevision).

There are no server processes here: every rank holds a replica and, for a table that has a logic attached, folds the
update batches of ALL workers of a clock in worker order through the same deterministic logic — each replica is "the
server", and they stay bit-identical.  Own writes are therefore not visible before the clock boundary on such tables
(as in the reference, where a client sees the server's result with its next row fetch).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


class AbstractServerTableLogic:
    def init(self, table) -> None:
        self.table = table

    def server_row_created(self, row_id: int, row) -> None:
        pass

    def apply_row_oplog(self, row_id: int, cols: torch.Tensor, updates: torch.Tensor, row, row_version: int,
                        end_of_version: bool) -> None:
        """Default server behaviour: the update IS the delta (server_table.cpp: ApplyRowOpLog -> BatchInc)."""
        row.apply_batch_inc(cols, updates)

    def server_row_sent(self, row_id: int, version: int, num_clients: int) -> None:
        pass

    def allow_send(self) -> bool:
        return True


class AdaRevisionServerTableLogic(AbstractServerTableLogic):
    """AdaRevision (McMahan & Streeter, "Delay-Tolerant Algorithms for Asynchronous Distributed Online Learning"): the
    workers push raw GRADIENTS; the server turns a gradient g that was computed on a stale copy of the row into

        g_bck   = accum_now - accum_at_the_version_the_worker_read        (gradients the worker has not seen)
        eta_old = alpha / sqrt(z_max);   z += g * (g + 2 * g_bck);   z_max = max(z, z_max);   eta = alpha / sqrt(z_max)
        delta   = -eta * g + (eta_old - eta) * g_bck;                accum += g

    per element (adarevision_server_table_logic.cpp:66-100 / 113-146), keeping one snapshot of ``accum`` per (row,
    version) that has been sent out and is still referenced by a client (ServerRowSent / end_of_version bookkeeping,
    :150-170)."""

    def __init__(self, init_step_size: float = 0.1, old_grad_upper_bound: int = 10000):
        self.alpha = float(init_step_size)
        self.upper = int(old_grad_upper_bound)
        self.info: Dict[int, dict] = {}
        self.old_accum: Dict[Tuple[int, int], list] = {}        # (row, version) -> [accum snapshot, clients still on it]

    def server_row_created(self, row_id, row):
        n, dev = row.capacity, row.device
        self.info[row_id] = {"accum": torch.zeros(n, device=dev), "z": torch.ones(n, device=dev),
                             "z_max": torch.ones(n, device=dev), "version": 0}

    def apply_row_oplog(self, row_id, cols, updates, row, row_version, end_of_version):
        st = self.info[row_id]
        g = torch.zeros(row.capacity, device=row.device)
        g.index_add_(0, torch.as_tensor(cols, dtype=torch.int64, device=row.device),
                     torch.as_tensor(updates, device=row.device).float())
        if row_version == 0:
            old = torch.zeros_like(g)
        else:
            key = (row_id, int(row_version))
            if key not in self.old_accum:
                raise KeyError(f"AdaRevision: no gradient snapshot for row {row_id} version {row_version}")
            old = self.old_accum[key][0]
        g_bck = st["accum"] - old
        eta_old = self.alpha / st["z_max"].sqrt()
        st["z"] += g * (g + 2 * g_bck)
        st["z_max"] = torch.maximum(st["z"], st["z_max"])
        eta = self.alpha / st["z_max"].sqrt()
        delta = -(eta * g) + (eta_old - eta) * g_bck
        st["accum"] += g
        if not torch.isfinite(delta).all():
            raise FloatingPointError(f"AdaRevision: non-finite delta in row {row_id}")
        row.apply_dense_batch_inc(delta.to(row.dtype))
        if row_version != 0 and end_of_version:
            ent = self.old_accum[(row_id, int(row_version))]
            ent[1] -= 1
            if ent[1] <= 0:
                del self.old_accum[(row_id, int(row_version))]

    def server_row_sent(self, row_id, version, num_clients):
        assert num_clients > 0
        self.old_accum[(row_id, int(version))] = [self.info[row_id]["accum"].clone(), int(num_clients)]

    def allow_send(self):
        return len(self.old_accum) < self.upper
