"""Write a LevelDB directory that stock LevelDB / Caffe / Poseidon can open (no libleveldb needed).

LevelDB is the reference's default database backend (tools/convert_imageset.cpp, tools/extract_features.cpp and
src/caffe/feature_extractor.cpp write it; src/caffe/layers/data_layer.cpp reads it).  A bulk load needs only the static
part of the format (leveldb/doc/{table_format,log_format,impl}.md):

    000005.ldb ...   sorted tables: 4 KiB data blocks (prefix-compressed entries, restart point every 16 keys, stored
                     uncompressed, trailer = type byte + masked CRC-32C), an empty metaindex block, the index block, the
                     48-byte footer with the table magic
    MANIFEST-000004  one log record holding a VersionEdit: comparator name, log / next-file numbers, last sequence and one
                     "new file" entry per table (level, number, size, smallest and largest internal key)
    CURRENT          "MANIFEST-000004\\n"
    000003.log       empty write-ahead log, LOCK, LOG

Tables are cut at ``table_bytes`` and registered at level 2 (sorted and non-overlapping, so any level works; level 0 would
trigger an immediate compaction when a writer opens the database).  Internal keys are ``user_key + fixed64(seq << 8 | 1)``
with sequence numbers 1..N.  Checksums use the C++ host module's CRC-32C.

Read back by the framework's own reader (csrc_host/leveldb_reader.cpp; tests/test_leveldb_writer.py); no libleveldb is
available here to cross-check.
"""
from __future__ import annotations

import os
import struct
from typing import Iterable, List, Tuple

_MAGIC = 0xDB4775248B80FB57
_BLOCK = 32768


def _varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _lp(b: bytes) -> bytes:
    return _varint(len(b)) + b


def _masked_crc(data: bytes) -> int:
    from . import native
    c = native.module().crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


class _BlockBuilder:
    def __init__(self, restart_interval: int = 16):
        self.ri = restart_interval
        self.buf = bytearray()
        self.restarts: List[int] = []
        self.count = 0
        self.last = b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count % self.ri == 0:
            self.restarts.append(len(self.buf))
        else:
            m = min(len(self.last), len(key))
            while shared < m and self.last[shared] == key[shared]:
                shared += 1
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def size(self) -> int:
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        restarts = self.restarts or [0]
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


class _TableWriter:
    def __init__(self, path: str, block_size: int):
        self.f = open(path, "wb")
        self.off = 0
        self.block_size = block_size
        self.data = _BlockBuilder()
        self.index = _BlockBuilder(restart_interval=1)
        self.smallest = self.largest = None

    def _write_block(self, raw: bytes) -> bytes:
        trailer = b"\0"               # kNoCompression
        crc = _masked_crc(raw + trailer)
        self.f.write(raw + trailer + struct.pack("<I", crc))
        handle = _varint(self.off) + _varint(len(raw))
        self.off += len(raw) + 5
        return handle

    def _flush(self):
        if self.data.count == 0:
            return
        last = self.data.last         # index key: any key >= the block's last and < the next block's first
        handle = self._write_block(self.data.finish())
        self.index.add(last, handle)
        self.data = _BlockBuilder()

    def add(self, ikey: bytes, value: bytes):
        if self.smallest is None:
            self.smallest = ikey
        self.largest = ikey
        self.data.add(ikey, value)
        if self.data.size() >= self.block_size:
            self._flush()

    def finish(self) -> int:
        self._flush()
        meta = self._write_block(_BlockBuilder().finish())
        index = self._write_block(self.index.finish())
        footer = meta + index
        footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
        self.f.write(footer)
        self.off += len(footer)
        self.f.close()
        return self.off


def _log_records(payloads: Iterable[bytes]) -> bytes:
    out = bytearray()
    for rec in payloads:
        pos, first = 0, True
        while True:
            left = _BLOCK - (len(out) % _BLOCK)
            if left < 7:
                out += b"\0" * left
                continue
            n = min(len(rec) - pos, left - 7)
            last = pos + n == len(rec)
            typ = 1 if (first and last) else 2 if first else 4 if last else 3
            frag = rec[pos: pos + n]
            out += struct.pack("<IHB", _masked_crc(bytes([typ]) + frag), n, typ) + frag
            pos += n
            first = False
            if last:
                break
    return bytes(out)


class LevelDBWriter:
    """``put`` / ``close`` front end (same interface as ``data.db.RecordWriter``): records are buffered, sorted by key on
    ``close`` — LevelDB keeps its keys ordered whatever the insertion order was — and bulk-loaded."""

    def __init__(self, path: str):
        self.path = path
        self.records = {}

    def put(self, key, value: bytes):
        self.records[key.encode() if isinstance(key, str) else bytes(key)] = bytes(value)     # last write wins

    def close(self):
        if self.records is not None:
            write_leveldb(self.path, sorted(self.records.items()))
            self.records = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def write_leveldb(path: str, records: Iterable[Tuple[bytes, bytes]], table_bytes: int = 64 << 20, block_size: int = 4096,
                  level: int = 2) -> int:
    """``records``: (key, value) pairs in strictly ascending key order (bytewise).  Returns the number written."""
    os.makedirs(path, exist_ok=True)
    for stale in os.listdir(path):
        if stale.endswith((".ldb", ".sst", ".log")) or stale.startswith("MANIFEST-") or stale in ("CURRENT", "LOCK", "LOG"):
            os.remove(os.path.join(path, stale))
    files = []                        # (number, size, smallest ikey, largest ikey)
    number = 5
    tw = None
    prev = None
    seq = 0
    for key, value in records:
        key, value = bytes(key), bytes(value)
        if prev is not None and not prev < key:
            raise ValueError("LevelDB writer: keys must be strictly ascending")
        prev = key
        seq += 1
        if tw is None:
            tw = _TableWriter(os.path.join(path, f"{number:06d}.ldb"), block_size)
        tw.add(key + struct.pack("<Q", (seq << 8) | 1), value)
        if tw.off + tw.data.size() >= table_bytes:
            files.append((number, tw.finish(), tw.smallest, tw.largest))
            number += 1
            tw = None
    if tw is not None:
        files.append((number, tw.finish(), tw.smallest, tw.largest))
        number += 1
    log_number = 3
    edit = _varint(1) + _lp(b"leveldb.BytewiseComparator") + _varint(2) + _varint(log_number) + \
        _varint(3) + _varint(number + 1) + _varint(4) + _varint(seq)
    for num, size, lo, hi in files:
        edit += _varint(7) + _varint(level) + _varint(num) + _varint(size) + _lp(lo) + _lp(hi)
    with open(os.path.join(path, "MANIFEST-000004"), "wb") as f:
        f.write(_log_records([edit]))
    with open(os.path.join(path, "CURRENT"), "w") as f:
        f.write("MANIFEST-000004\n")
    open(os.path.join(path, f"{log_number:06d}.log"), "wb").close()
    open(os.path.join(path, "LOCK"), "wb").close()
    with open(os.path.join(path, "LOG"), "w") as f:
        f.write(f"poseidon_b200.data.leveldb_writer: bulk load of {seq} records into {len(files)} table(s) at level {level}\n")
    return seq
