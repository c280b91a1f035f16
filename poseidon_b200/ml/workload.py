"""Which samples does worker (client, thread) visit, in which mini-batches (reference: ps/src/ml/util/
workload_manager.hpp:12-125).

``global_data``: every client sees the whole dataset, so it is split over clients x threads; otherwise the dataset is
already this client's share and is split over its threads only (the last thread takes the remainder).  An epoch is
``num_batches_per_epoch`` batches of ``ceil(n / num_batches)`` samples, wrapping around inside the worker's range when
that does not divide.

The global split here is the balanced one (the first ``n mod T`` workers get one sample more); the reference's formula
for workers past its cut-off overlaps the neighbouring range by one sample (workload_manager.hpp:47-50)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List


@dataclass
class WorkloadManagerConfig:
    thread_id: int = 0
    client_id: int = 0
    num_clients: int = 1
    num_threads: int = 1
    num_batches_per_epoch: int = 1
    num_data: int = 0
    global_data: bool = True


class WorkloadManager:
    def __init__(self, config: WorkloadManagerConfig):
        c = config
        n = c.num_data
        if c.global_data:
            total = c.num_clients * c.num_threads
            me = c.client_id * c.num_threads + c.thread_id
            q, r = divmod(n, total)
            self.begin = me * q + min(me, r)
            self.end = self.begin + q + (1 if me < r else 0)
        else:
            per = n // c.num_threads
            self.begin = per * c.thread_id
            self.end = n if c.thread_id == c.num_threads - 1 else self.begin + per
        span = self.end - self.begin
        self.batch_size = math.ceil(span / c.num_batches_per_epoch) if span > 0 else 0
        if self.batch_size <= 0:
            raise ValueError(f"batch size cannot be 0: {span} samples for this worker, "
                             f"{c.num_batches_per_epoch} batches per epoch")
        self.num_data_per_epoch = self.batch_size * c.num_batches_per_epoch
        self.restart()

    def get_batch_size(self) -> int:
        return self.batch_size

    def get_num_batches(self) -> int:
        return self.num_data_per_epoch // self.batch_size

    def restart(self):
        self.seen = 0

    def _wrap(self, idx: int) -> int:
        return idx if idx < self.end else (idx - self.end) % (self.end - self.begin) + self.begin

    def get_data_idx_and_advance(self) -> int:
        if self.is_end():
            raise RuntimeError("end of epoch: call restart()")
        idx = self._wrap(self.begin + self.seen)
        self.seen += 1
        return idx

    def get_batch_data_idx(self, num_data: int) -> List[int]:
        """The next ``num_data`` indices, without advancing."""
        return [self._wrap(self.begin + self.seen + i) for i in range(num_data)]

    def is_end(self) -> bool:
        return self.seen == self.num_data_per_epoch

    def is_end_of_batch(self) -> bool:
        return self.seen % self.batch_size == 0
