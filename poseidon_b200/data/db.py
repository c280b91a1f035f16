"""Record databases holding serialized ``Datum`` protos.

The reference reads LevelDB/LMDB (src/caffe/layers/data_layer.cpp:102-140); neither
library exists in this image, so the framework ships its own flat-file record store
("PDB") with the same cursor semantics (ordered keys, seek-to-first, next, wrap), reads
existing LMDB environments with a built-in parser of the ``data.mdb`` format
(``lmdb_reader.py``; the ``lmdb`` module is used instead when importable) and existing
LevelDB directories through the C++ host runtime (``csrc_host/leveldb_reader.cpp``).

PDB layout:  b"PDB1" | u64 n | n × (u32 klen, u32 vlen, key, value)   — little endian.
"""
from __future__ import annotations

import os
import struct
from typing import Iterator, List, Tuple

from .. import proto as P

_MAGIC = b"PDB1"


class RecordWriter:
    def __init__(self, path: str):
        self.path = path
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        self.f = open(path, "wb")
        self.f.write(_MAGIC + struct.pack("<Q", 0))
        self.n = 0

    def put(self, key, value: bytes):
        key = key.encode() if isinstance(key, str) else key
        self.f.write(struct.pack("<II", len(key), len(value)))
        self.f.write(key)
        self.f.write(value)
        self.n += 1

    def close(self):
        self.f.seek(len(_MAGIC))
        self.f.write(struct.pack("<Q", self.n))
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class RecordReader:
    """Whole-file index (offsets only); values are read lazily via ``os.pread``."""

    def __init__(self, path: str):
        if os.path.isdir(path):
            cand = os.path.join(path, "data.pdb")
            if os.path.exists(cand):
                path = cand
        self.path = path
        self.fd = os.open(path, os.O_RDONLY)
        head = os.pread(self.fd, 12, 0)
        if head[:4] != _MAGIC:
            os.close(self.fd)
            raise IOError(f"{path}: not a PDB record file")
        (self.n,) = struct.unpack("<Q", head[4:12])
        self.index: List[Tuple[int, int, int]] = []   # (key_off, klen, vlen)
        off = 12
        size = os.fstat(self.fd).st_size
        while off < size and len(self.index) < self.n:
            klen, vlen = struct.unpack("<II", os.pread(self.fd, 8, off))
            self.index.append((off + 8, klen, vlen))
            off += 8 + klen + vlen

    def __len__(self):
        return len(self.index)

    def key(self, i: int) -> bytes:
        off, klen, _ = self.index[i]
        return os.pread(self.fd, klen, off)

    def value(self, i: int) -> bytes:
        off, klen, vlen = self.index[i]
        return os.pread(self.fd, vlen, off + klen)

    def datum(self, i: int):
        return P.Datum.FromString(self.value(i))

    def __iter__(self) -> Iterator[Tuple[bytes, bytes]]:
        for i in range(len(self)):
            yield self.key(i), self.value(i)

    def close(self):
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None


class LMDBReader:
    """Thin adaptor (used only when the optional ``lmdb`` module is importable)."""

    def __init__(self, path: str):
        import lmdb  # noqa: F401  (optional)
        self.env = lmdb.open(path, readonly=True, lock=False, readahead=True)
        with self.env.begin() as txn:
            self.keys = [bytes(k) for k, _ in txn.cursor()]

    def __len__(self):
        return len(self.keys)

    def key(self, i):
        return self.keys[i]

    def value(self, i):
        with self.env.begin() as txn:
            return bytes(txn.get(self.keys[i]))

    def datum(self, i):
        return P.Datum.FromString(self.value(i))

    def close(self):
        self.env.close()


def open_db(path: str, backend: str = "LEVELDB"):
    """Open ``path`` as PDB (file or dir/data.pdb), falling back to LMDB if available."""
    pdb = path if os.path.isfile(path) else os.path.join(path, "data.pdb")
    if os.path.isfile(pdb):
        return RecordReader(pdb)
    if os.path.isdir(path) and os.path.isfile(os.path.join(path, "data.mdb")):
        try:
            return LMDBReader(path)                     # liblmdb bindings, when installed
        except ImportError:
            from .lmdb_reader import LMDBFile
            return LMDBFile(path)                       # built-in read-only parser of the data.mdb format
    if os.path.isdir(path) and os.path.isfile(os.path.join(path, "CURRENT")):
        from . import native
        if native.available():
            return native.NativeRecordDB(path)          # LevelDB directory, read by the C++ host runtime
    raise IOError(f"cannot open database '{path}' (backend {backend}): no PDB store, LMDB environment or LevelDB "
                  "directory found there — create one with tools.convert_imageset, or run with synthetic data")


def shard_indices(n_records: int, shared_fs: bool, num_clients: int, client_id: int,
                  num_threads: int, thread_id: int) -> Tuple[int, int]:
    """(offset, stride) of the records one worker reads.

    shared file system: all workers read the same DB, worker ``client*threads+thread`` takes
    every ``clients*threads``-th record; otherwise each client opens ``source_<client_id>``
    and its threads stride by ``num_threads``.
    reference: src/caffe/layers/data_layer.cpp:143-161, image_data_layer.cpp:36-55."""
    if shared_fs:
        return client_id * num_threads + thread_id, num_clients * num_threads
    return thread_id, num_threads
