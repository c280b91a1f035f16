"""Lightweight always-on statistics: named timers / counters / byte tallies, dumped as YAML
(``--stats_path.<rank>``) at shutdown — the role of Bösen's PETUUM_STATS macros.

reference: ps/src/petuum_ps_common/util/stats.hpp:17-456 (STATS_* macros), :537-773
(per-thread structs), stats.cpp:1312-1844 (YAML dump at ~TableGroup).
"""
from __future__ import annotations

import contextlib
import threading
import time
from collections import defaultdict


class Stats:
    def __init__(self):
        self.lock = threading.Lock()
        self.reset()

    def reset(self):
        self.timers = defaultdict(float)
        self.timer_calls = defaultdict(int)
        self.counters = defaultdict(int)
        self.values = {}

    @contextlib.contextmanager
    def timer(self, name: str):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            dt = time.perf_counter() - t0
            with self.lock:
                self.timers[name] += dt
                self.timer_calls[name] += 1

    def count(self, name: str, n: int = 1):
        with self.lock:
            self.counters[name] += n

    def set(self, name: str, value):
        with self.lock:
            self.values[name] = value

    def as_dict(self):
        with self.lock:
            return {
                "timers_sec": dict(self.timers),
                "timer_calls": dict(self.timer_calls),
                "counters": dict(self.counters),
                "values": dict(self.values),
            }

    def dump_yaml(self, path: str):
        import yaml
        with open(path, "w") as f:
            yaml.safe_dump(self.as_dict(), f, default_flow_style=False)


STATS = Stats()
