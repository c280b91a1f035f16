"""Process-wide flag store — every command-line flag snapshotted as a string and readable anywhere by name —
plus the SVB on/off/completed switches the reference keeps there.

reference: include/caffe/context.hpp:17-101, src/caffe/context.cpp:17-29 (gflags snapshot), :44-53 (InitSVB),
:57-74 (get_int32/get_double/get_bool/get_string), :96-123 (parse_int_list).
"""
from __future__ import annotations

import threading
from typing import Dict, List


class Context:
    _instance = None
    _lock = threading.Lock()

    def __init__(self):
        self._flags: Dict[str, str] = {}
        self.use_svb = False
        self.svb_completed = False

    @classmethod
    def get_instance(cls) -> "Context":
        with cls._lock:
            if cls._instance is None:
                cls._instance = Context()
            return cls._instance

    def load(self, namespace) -> "Context":
        """Snapshot an argparse namespace (or dict) as strings."""
        items = vars(namespace) if not isinstance(namespace, dict) else namespace
        for k, v in items.items():
            if v is not None:
                self._flags[k] = str(v)
        if "num_table_threads" in self._flags:
            self._flags["num_app_threads"] = str(max(1, int(self._flags["num_table_threads"]) - 1))
        return self

    def set(self, key, value):
        self._flags[key] = str(value)

    def _get(self, key):
        if key not in self._flags:
            raise KeyError(f"Failed to lookup {key} in params")
        return self._flags[key]

    def get_int32(self, key) -> int:
        return int(float(self._get(key)))

    get_int64 = get_int32

    def get_double(self, key) -> float:
        return float(self._get(key))

    def get_bool(self, key) -> bool:
        return self._get(key).lower() in ("1", "true", "yes", "on")

    def get_string(self, key) -> str:
        return self._get(key)

    @staticmethod
    def parse_int_list(s: str, sep: str = ",") -> List[int]:
        return [int(x) for x in s.split(sep) if x.strip() != ""]

    def set_use_svb(self, v: bool):
        self.use_svb = bool(v)

    def set_svb_completed(self, v: bool = True):
        self.svb_completed = bool(v)
