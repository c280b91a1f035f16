"""``key: value`` meta files that accompany PMLS datasets (reference: ps/src/ml/util/metafile_reader.{hpp,cpp})."""
from __future__ import annotations


class MetafileReader:
    def __init__(self, path: str = None):
        self.content = {}
        if path:
            self.init(path)

    def init(self, path: str):
        self.path = path
        with open(path) as f:
            for line in f:
                line = line.strip()
                if not line or line.startswith("#") or ":" not in line:
                    continue
                k, v = line.split(":", 1)
                self.content[k.strip()] = v.strip()

    def _get(self, key: str) -> str:
        if key not in self.content:
            raise KeyError(f"{key} not found in metafile {getattr(self, 'path', '?')}")
        return self.content[key]

    def get_int32(self, key: str) -> int:
        return int(self._get(key))

    def get_double(self, key: str) -> float:
        return float(self._get(key))

    def get_bool(self, key: str) -> bool:
        return self._get(key).lower() in ("1", "true", "yes")

    def get_string(self, key: str) -> str:
        return self._get(key)
