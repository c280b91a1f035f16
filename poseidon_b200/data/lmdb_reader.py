"""Read-only LMDB (data.mdb) parser — no liblmdb needed.

Caffe / Poseidon datasets are LMDB (or LevelDB) environments of ``key -> Datum`` records
(reference: src/caffe/layers/data_layer.cpp:102-140 opens them with ``MDB_RDONLY|MDB_NOTLS`` and walks a cursor with
``MDB_FIRST`` / ``MDB_NEXT``).  This module implements exactly that read path from the on-disk format of LMDB 0.9
(``mdb.c``): two meta pages, a B+tree of branch / leaf pages, overflow pages for values larger than a page.

    page header (16 B) : pgno u64 | pad u16 | flags u16 | lower u16, upper u16   (overflow pages: page-count u32)
    meta               : magic 0xBEEFC0DE u32 | version u32 | address u64 | mapsize u64 | MDB_db free | MDB_db main |
                         last_pg u64 | txnid u64                (MDB_db = pad u32, flags u16, depth u16, 3 x u64 page
                         counts, entries u64, root u64; the free DB's ``pad`` field holds the page size)
    node               : lo u16 | hi u16 | flags u16 | ksize u16 | key | (leaf) data   — branch: child pgno =
                         lo | hi<<16 | flags<<32 ; leaf: data size = lo | hi<<16, F_BIGDATA (0x01): data is the u64
                         page number of an overflow run whose payload starts after its 16-byte header

Only what a plain (non-DUPSORT) database needs is implemented; sub-databases raise.  There is no liblmdb in the build
image, so the parser is validated against files produced by the test-suite's own writer, which follows the same
definition (tests/test_lmdb_reader.py).
"""
from __future__ import annotations

import mmap
import os
import struct
from typing import Iterator, List, Tuple

P_BRANCH, P_LEAF, P_OVERFLOW, P_META, P_LEAF2 = 0x01, 0x02, 0x04, 0x08, 0x20
F_BIGDATA, F_SUBDATA, F_DUPDATA = 0x01, 0x02, 0x04
MDB_MAGIC = 0xBEEFC0DE
PAGEHDRSZ = 16
_INVALID = 0xFFFFFFFFFFFFFFFF


class LMDBFormatError(IOError):
    pass


class LMDBFile:
    """Ordered read-only view of the main database of an LMDB environment."""

    def __init__(self, path: str):
        mdb = os.path.join(path, "data.mdb") if os.path.isdir(path) else path
        self.path = mdb
        self.f = open(mdb, "rb")
        size = os.fstat(self.f.fileno()).st_size
        if size < 2 * 512:
            raise LMDBFormatError(f"{mdb}: too small to be an LMDB file")
        self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        meta0 = self._read_meta(0)
        self.psize = meta0["psize"]
        if self.psize < 512 or self.psize > 65536 or self.psize & (self.psize - 1):
            raise LMDBFormatError(f"{mdb}: implausible page size {self.psize}")
        meta1 = self._read_meta(self.psize)
        meta = meta1 if meta1["txnid"] > meta0["txnid"] else meta0
        self.entries = meta["entries"]
        self.root = meta["root"]
        self.depth = meta["depth"]
        if meta["flags"] & 0x04:          # MDB_DUPSORT
            raise LMDBFormatError(f"{mdb}: DUPSORT databases are not supported")
        self._index: List[Tuple[int, int]] | None = None      # (leaf page offset, node offset) per record, built lazily

    # ------------------------------------------------------------------------------------------------ low level
    def _read_meta(self, off: int):
        pgno, pad, flags = struct.unpack_from("<QHH", self.mm, off)
        if not flags & P_META:
            raise LMDBFormatError(f"{self.path}: page at {off} is not a meta page")
        magic, version, _addr, _mapsize = struct.unpack_from("<IIQQ", self.mm, off + PAGEHDRSZ)
        if magic != MDB_MAGIC:
            raise LMDBFormatError(f"{self.path}: bad magic {magic:#x}")
        if version != 1:
            raise LMDBFormatError(f"{self.path}: unsupported LMDB data version {version}")
        dbs = off + PAGEHDRSZ + 24
        free_pad, _ff, _fd = struct.unpack_from("<IHH", self.mm, dbs)
        mpad, mflags, mdepth, _b, _l, _o, mentries, mroot = struct.unpack_from("<IHHQQQQQ", self.mm, dbs + 48)
        _last, txnid = struct.unpack_from("<QQ", self.mm, dbs + 96)
        return {"psize": free_pad, "flags": mflags, "depth": mdepth, "entries": mentries, "root": mroot, "txnid": txnid}

    def _page(self, pgno: int) -> int:
        off = pgno * self.psize
        if off + PAGEHDRSZ > len(self.mm):
            raise LMDBFormatError(f"{self.path}: page {pgno} beyond end of file")
        return off

    def _nodes(self, off: int):
        flags, lower = struct.unpack_from("<HH", self.mm, off + 10)
        n = (lower - PAGEHDRSZ) >> 1
        ptrs = struct.unpack_from(f"<{n}H", self.mm, off + PAGEHDRSZ) if n > 0 else ()
        return flags, ptrs

    def _leaf_record(self, off: int, ptr: int) -> Tuple[bytes, bytes]:
        lo, hi, nflags, ksize = struct.unpack_from("<HHHH", self.mm, off + ptr)
        kstart = off + ptr + 8
        key = bytes(self.mm[kstart:kstart + ksize])
        dsize = lo | (hi << 16)
        if nflags & (F_SUBDATA | F_DUPDATA):
            raise LMDBFormatError(f"{self.path}: sub-databases / duplicates are not supported")
        if nflags & F_BIGDATA:
            (opg,) = struct.unpack_from("<Q", self.mm, kstart + ksize)
            ooff = self._page(opg)
            oflags = struct.unpack_from("<H", self.mm, ooff + 10)[0]
            if not oflags & P_OVERFLOW:
                raise LMDBFormatError(f"{self.path}: page {opg} is not an overflow page")
            data = bytes(self.mm[ooff + PAGEHDRSZ: ooff + PAGEHDRSZ + dsize])
        else:
            data = bytes(self.mm[kstart + ksize: kstart + ksize + dsize])
        return key, data

    def _walk(self, pgno: int, depth: int = 0) -> Iterator[Tuple[int, int]]:
        """In-order traversal yielding (leaf page offset, node offset)."""
        if depth > 64:
            raise LMDBFormatError(f"{self.path}: tree too deep (corrupt?)")
        off = self._page(pgno)
        flags, ptrs = self._nodes(off)
        if flags & P_LEAF:
            if flags & P_LEAF2:
                raise LMDBFormatError(f"{self.path}: LEAF2 pages (DUPFIXED) are not supported")
            for p in ptrs:
                yield off, p
        elif flags & P_BRANCH:
            for p in ptrs:
                lo, hi, nflags = struct.unpack_from("<HHH", self.mm, off + p)
                child = lo | (hi << 16) | (nflags << 32)
                yield from self._walk(child, depth + 1)
        else:
            raise LMDBFormatError(f"{self.path}: unexpected page flags {flags:#x} at page {pgno}")

    # ------------------------------------------------------------------------------------------------ reader API
    def _build_index(self):
        if self._index is None:
            self._index = [] if self.root == _INVALID else list(self._walk(self.root))
            if self.entries and len(self._index) != self.entries:
                raise LMDBFormatError(f"{self.path}: walked {len(self._index)} records, meta says {self.entries}")

    def __len__(self):
        self._build_index()
        return len(self._index)

    def key(self, i: int) -> bytes:
        self._build_index()
        return self._leaf_record(*self._index[i])[0]

    def value(self, i: int) -> bytes:
        self._build_index()
        return self._leaf_record(*self._index[i])[1]

    def datum(self, i: int):
        from .. import proto as P
        return P.Datum.FromString(self.value(i))

    def __iter__(self):
        self._build_index()
        for ref in self._index:
            yield self._leaf_record(*ref)

    def close(self):
        if self.mm is not None:
            self.mm.close()
            self.f.close()
            self.mm = None
