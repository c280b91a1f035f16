"""Dependency-free HDF5 reader / writer for the HDF5_DATA and HDF5_OUTPUT layers (no libhdf5 / h5py in the image).

reference: src/caffe/layers/hdf5_data_layer.cpp:38-108 (loads whole "data" / "label" datasets of every listed file into
memory through H5LT), src/caffe/layers/hdf5_output_layer.cpp, src/caffe/util/io.cpp (hdf5_load_nd_dataset: float or double
datasets of 1..4 dimensions).

What is implemented is the part of the HDF5 file format that such files use:

* superblock versions 0 and 1 (what libhdf5 writes unless ``libver='latest'`` is requested), optional user block;
* "old style" groups: symbol-table message -> v1 B-tree (``TREE``) -> symbol nodes (``SNOD``) + local heap (``HEAP``),
  nested groups included;
* version-1 object headers with continuation blocks;
* dataspace messages v1 / v2 (simple dataspaces), datatype classes 0 (integers) and 1 (IEEE floats), little or big endian;
* data layout message v3 (compact, contiguous, chunked) and the v1 / v2 layouts of very old files;
* chunked storage through the v1 chunk B-tree, with the deflate (gzip) and shuffle filters — ``h5py``'s
  ``compression="gzip"``, ``shuffle=True``.

Not implemented (a clear error names the feature): superblock v2 / v3 with new-style groups (fractal heaps, ``OHDR``
version-2 headers), variable-length / compound / string datatypes, szip / lzf / scale-offset filters, external storage,
virtual datasets.

The reader is validated against a file written by a genuine HDF5 library (the MATLAB v7.3 sample that ships with scipy,
tests/test_hdf5.py); the writer emits the same structures (superblock 0, symbol-table root group, v1 headers, contiguous
little-endian float / double / integer datasets) and is validated against the reader.
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(IOError):
    pass


# ======================================================================================================== reader
class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.base = 0
        self.O = 8          # size of offsets
        self.L = 8          # size of lengths

    def u(self, off: int, n: int) -> int:
        return int.from_bytes(self.buf[off: off + n], "little")

    def addr(self, off: int) -> int:
        """File address stored at ``off`` -> absolute position in the buffer (UNDEF stays UNDEF)."""
        v = self.u(off, self.O)
        if v == (1 << (8 * self.O)) - 1:
            return UNDEF
        return v + self.base


class Dataset:
    def __init__(self, name: str):
        self.name = name
        self.shape: Tuple[int, ...] = ()
        self.dtype: Optional[np.dtype] = None
        self.layout = None           # ("compact", bytes) | ("contiguous", addr, size) | ("chunked", btree, chunk_dims)
        self.filters: List[Tuple[int, Tuple[int, ...]]] = []

    def __repr__(self):
        return f"<HDF5 dataset {self.name!r} shape={self.shape} dtype={self.dtype}>"


class File:
    """Read-only view of an HDF5 file: ``f["data"]`` -> numpy array, ``f.keys()``, nested groups as ``"g/name"``."""

    _CORRUPT = (IndexError, ValueError, struct.error, zlib.error, OverflowError, MemoryError, RecursionError)

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            buf = fh.read()
        self.path = path
        try:
            self._parse(buf)
        except HDF5Error:
            raise
        except self._CORRUPT as e:
            raise HDF5Error(f"{path}: corrupt or truncated HDF5 file ({type(e).__name__}: {e})") from e

    def _parse(self, buf: bytes):
        path = self.path
        r = self.r = _Reader(buf)
        sb = -1
        off = 0
        while off + 8 <= len(buf):                       # the superblock sits at 0, 512, 1024, 2048, ... (user block)
            if buf[off: off + 8] == SIGNATURE:
                sb = off
                break
            off = 512 if off == 0 else off * 2
        if sb < 0:
            raise HDF5Error(f"{path}: not an HDF5 file (signature not found)")
        ver = buf[sb + 8]
        if ver > 1:
            raise HDF5Error(f"{path}: superblock version {ver} (libver='latest' files with new-style groups) is not "
                            "supported; rewrite the file with the default (earliest) format")
        r.O, r.L = buf[sb + 13], buf[sb + 14]
        p = sb + 24 + (4 if ver == 1 else 0)
        r.base = 0
        base = r.u(p, r.O)
        r.base = base
        p += 4 * r.O                                        # base, free-space, end-of-file, driver-info addresses
        # root group symbol table entry
        self.datasets: Dict[str, Dataset] = {}
        root_hdr = r.addr(p + r.O)
        cache_type = r.u(p + 2 * r.O, 4)
        if cache_type == 1:
            btree, heap = r.addr(p + 2 * r.O + 8), r.addr(p + 2 * r.O + 8 + r.O)
            self._walk_group(btree, heap, "")
        else:
            self._visit_object(root_hdr, "")

    # ------------------------------------------------------------------------------------------------ groups
    def _heap_data(self, heap: int) -> int:
        r = self.r
        if r.buf[heap: heap + 4] != b"HEAP":
            raise HDF5Error("local heap signature missing")
        return r.addr(heap + 8 + 2 * r.L)

    def _walk_group(self, btree: int, heap: int, prefix: str):
        r = self.r
        names_at = self._heap_data(heap)
        for snod in self._group_leaves(btree):
            if r.buf[snod: snod + 4] != b"SNOD":
                raise HDF5Error("symbol node signature missing")
            n = r.u(snod + 6, 2)
            esz = 2 * r.O + 24
            for i in range(n):
                e = snod + 8 + i * esz
                name_off = r.u(e, r.O)
                end = r.buf.index(b"\0", names_at + name_off)
                name = r.buf[names_at + name_off: end].decode("utf-8", "replace")
                hdr = r.addr(e + r.O)
                self._visit_object(hdr, prefix + name)

    def _group_leaves(self, node: int):
        r = self.r
        if r.buf[node: node + 4] != b"TREE":
            raise HDF5Error("B-tree signature missing")
        ntype, level, used = r.buf[node + 4], r.buf[node + 5], r.u(node + 6, 2)
        if ntype != 0:
            raise HDF5Error("expected a group B-tree")
        p = node + 8 + 2 * r.O
        for i in range(used):
            child = r.addr(p + r.L + i * (r.L + r.O))
            if level == 0:
                yield child
            else:
                yield from self._group_leaves(child)

    # ------------------------------------------------------------------------------------------------ objects
    def _messages(self, hdr: int):
        """(type, flags, absolute offset of the message body, size) for a version-1 object header + continuations."""
        r = self.r
        if r.buf[hdr: hdr + 4] == b"OHDR":
            raise HDF5Error("version-2 object headers (libver='latest') are not supported")
        if r.buf[hdr] != 1:
            raise HDF5Error(f"object header version {r.buf[hdr]} not supported")
        nmsg = r.u(hdr + 2, 2)
        size = r.u(hdr + 8, 4)
        blocks = [(hdr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = r.u(p, 2), r.u(p + 2, 2), r.buf[p + 4]
                body = p + 8
                if mtype == 0x10:                            # continuation
                    blocks.append((r.addr(body), r.u(body + r.O, r.L)))
                out.append((mtype, flags, body, msize))
                p = body + msize
        return out

    def _visit_object(self, hdr: int, name: str):
        r = self.r
        msgs = self._messages(hdr)
        types = {m[0] for m in msgs}
        if 0x11 in types:                                    # a group: symbol table message
            body = next(m for m in msgs if m[0] == 0x11)[2]
            self._walk_group(r.addr(body), r.addr(body + r.O), name + "/" if name else "")
            return
        if 0x08 not in types or 0x01 not in types or 0x03 not in types:
            if 0x02 in types or 0x06 in types:
                raise HDF5Error("new-style groups (link messages) are not supported")
            return                                           # named datatype or something else without data
        ds = Dataset(name)
        for mtype, flags, body, msize in msgs:
            if mtype == 0x01:
                ds.shape = self._dataspace(body)
            elif mtype == 0x03:
                try:
                    ds.dtype = self._datatype(body)
                except HDF5Error:
                    ds.dtype = None                          # unsupported type: listed, error on access
            elif mtype == 0x08:
                ds.layout = self._layout(body)
            elif mtype == 0x0B:
                ds.filters = self._filters(body)
        self.datasets[name] = ds

    def _dataspace(self, p: int) -> Tuple[int, ...]:
        r = self.r
        ver, rank = r.buf[p], r.buf[p + 1]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if r.buf[p + 3] == 2:
                raise HDF5Error("null dataspace")
            q = p + 4
        else:
            raise HDF5Error(f"dataspace message version {ver}")
        return tuple(r.u(q + i * r.L, r.L) for i in range(rank))

    def _datatype(self, p: int) -> np.dtype:
        r = self.r
        cls, bits0, size = r.buf[p] & 0x0F, r.buf[p + 1], r.u(p + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            kind = "i" if bits0 & 0x08 else "u"
            if size not in (1, 2, 4, 8):
                raise HDF5Error(f"integer size {size}")
            return np.dtype(f"{order}{kind}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise HDF5Error(f"float size {size}")
            return np.dtype(f"{order}f{size}")
        raise HDF5Error(f"datatype class {cls} (only integers and IEEE floats are supported)")

    def _layout(self, p: int):
        r = self.r
        ver = r.buf[p]
        if ver == 3:
            cls = r.buf[p + 1]
            if cls == 0:
                n = r.u(p + 2, 2)
                return ("compact", bytes(r.buf[p + 4: p + 4 + n]))
            if cls == 1:
                return ("contiguous", r.addr(p + 2), r.u(p + 2 + r.O, r.L))
            if cls == 2:
                nd = r.buf[p + 2]
                bt = r.addr(p + 3)
                dims = tuple(r.u(p + 3 + r.O + 4 * i, 4) for i in range(nd))
                return ("chunked", bt, dims)
            raise HDF5Error(f"layout class {cls}")
        if ver in (1, 2):
            nd, cls = r.buf[p + 1], r.buf[p + 2]
            q = p + 8
            a = None
            if cls != 0:
                a = r.addr(q)
                q += r.O
            dims = tuple(r.u(q + 4 * i, 4) for i in range(nd))
            q += 4 * nd
            if cls == 0:
                n = r.u(q, 4)
                return ("compact", bytes(r.buf[q + 4: q + 4 + n]))
            if cls == 1:
                return ("contiguous", a, None)
            return ("chunked", a, dims)
        raise HDF5Error(f"data layout message version {ver} (libver='latest') is not supported")

    def _filters(self, p: int):
        r = self.r
        ver, n = r.buf[p], r.buf[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = r.u(q, 2)
            if ver == 1 or fid >= 256:
                nlen = r.u(q + 2, 2)
                q += 4
            else:
                nlen = 0
                q += 2
            ncd = r.u(q + 2, 2)
            q += 4
            if ver == 1:
                nlen = (nlen + 7) // 8 * 8
            q += nlen
            cd = tuple(r.u(q + 4 * i, 4) for i in range(ncd))
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    # ------------------------------------------------------------------------------------------------ data
    def keys(self):
        return list(self.datasets)

    def __contains__(self, name):
        return name.strip("/") in self.datasets

    def __getitem__(self, name: str) -> np.ndarray:
        ds = self.datasets.get(name.strip("/"))
        if ds is None:
            raise KeyError(f"{self.path}: no dataset '{name}' (have: {', '.join(self.datasets) or 'none'})")
        return self.read(ds)

    def read(self, ds: Dataset) -> np.ndarray:
        try:
            return self._read(ds)
        except HDF5Error:
            raise
        except self._CORRUPT as e:
            raise HDF5Error(f"{self.path}: dataset '{ds.name}' is corrupt or truncated ({type(e).__name__}: {e})") from e

    def _read(self, ds: Dataset) -> np.ndarray:
        r = self.r
        if ds.dtype is None:
            raise HDF5Error(f"dataset '{ds.name}': unsupported datatype")
        count = 1
        for d in ds.shape:
            count *= int(d)
        if count * ds.dtype.itemsize > (1 << 40):
            raise HDF5Error(f"dataset '{ds.name}': implausible extent {ds.shape}")
        nbytes = count * ds.dtype.itemsize
        kind = ds.layout[0]
        if kind == "compact":
            raw = ds.layout[1][:nbytes]
        elif kind == "contiguous":
            a = ds.layout[1]
            if a == UNDEF:                                   # never written: fill value (zeros)
                return np.zeros(ds.shape, ds.dtype.newbyteorder("="))
            raw = r.buf[a: a + nbytes]
        else:
            return self._read_chunked(ds)
        if len(raw) < nbytes:
            raise HDF5Error(f"dataset '{ds.name}': file truncated")
        return np.frombuffer(raw, ds.dtype, count).reshape(ds.shape).astype(ds.dtype.newbyteorder("="))

    def _chunks(self, node: int, nd: int):
        r = self.r
        if node == UNDEF:
            return
        if r.buf[node: node + 4] != b"TREE" or r.buf[node + 4] != 1:
            raise HDF5Error("chunk B-tree signature missing")
        level, used = r.buf[node + 5], r.u(node + 6, 2)
        p = node + 8 + 2 * r.O
        ksz = 8 + 8 * nd
        for i in range(used):
            k = p + i * (ksz + r.O)
            child = r.addr(k + ksz)
            if level == 0:
                yield (r.u(k, 4), r.u(k + 4, 4), tuple(r.u(k + 8 + 8 * j, 8) for j in range(nd - 1)), child)
            else:
                yield from self._chunks(child, nd)

    def _read_chunked(self, ds: Dataset) -> np.ndarray:
        r = self.r
        _, bt, cdims = ds.layout
        nd = len(cdims)                                      # rank + 1 (last = element size)
        cshape = cdims[:-1]
        esz = ds.dtype.itemsize
        out = np.zeros(ds.shape, ds.dtype)
        for size, mask, offs, a in self._chunks(bt, nd):
            raw = bytes(r.buf[a: a + size])
            for idx in range(len(ds.filters) - 1, -1, -1):   # undo the pipeline, last filter first
                if mask & (1 << idx):
                    continue
                fid, cd = ds.filters[idx]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    n = len(raw) // esz
                    raw = np.frombuffer(raw, np.uint8, n * esz).reshape(esz, n).T.tobytes() + raw[n * esz:]
                elif fid == 3:
                    raw = raw[:-4]                           # fletcher32 checksum trailer
                else:
                    raise HDF5Error(f"dataset '{ds.name}': filter {fid} is not supported (gzip / shuffle / fletcher32 are)")
            chunk = np.frombuffer(raw, ds.dtype, int(np.prod(cshape))).reshape(cshape)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, ds.shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]
        return out.astype(ds.dtype.newbyteorder("="))

    def close(self):
        self.r = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def load(path: str, names=("data", "label")) -> Dict[str, np.ndarray]:
    with File(path) as f:
        return {n: f[n] for n in names}


# ======================================================================================================== writer
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _count_msgs(msgs: bytes) -> int:
    n = p = 0
    while p < len(msgs):
        p += 8 + struct.unpack_from("<H", msgs, p + 2)[0]
        n += 1
    return n


class _Window:
    """Write-only window into a growing bytearray (a memoryview would pin its size)."""

    def __init__(self, buf: bytearray, off: int):
        self.buf, self.off = buf, off

    def __setitem__(self, sl, value):
        start = sl.start or 0
        self.buf[self.off + start: self.off + start + len(value)] = value


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 4:
            return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBIHH", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    raise HDF5Error(f"cannot write dtype {dt}")


_CHUNK_K = 32            # "indexed storage internal node K": the library's default, implied by superblock version 0


def _shuffle(raw: bytes, esz: int) -> bytes:
    n = len(raw) // esz
    return np.frombuffer(raw, np.uint8, n * esz).reshape(n, esz).T.tobytes() + raw[n * esz:]


def _chunk_tree(entries, nd: int, alloc):
    """v1 B-tree of type 1 over ``entries`` = [(nbytes, offsets, address)] in row-major chunk order.  ``alloc(nbytes)``
    returns (address, bytearray view to fill).  Returns the root address."""
    ksz = 8 + 8 * (nd + 1)
    node_size = 24 + (2 * _CHUNK_K + 1) * ksz + 2 * _CHUNK_K * 8

    def key(nbytes, offs):
        return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

    level = 0
    items = [(key(n, o), a, o) for n, o, a in entries]          # (first key, child address, first offsets)
    last_offs = entries[-1][1]
    while True:
        groups = [items[i: i + 2 * _CHUNK_K] for i in range(0, len(items), 2 * _CHUNK_K)]
        addrs = [alloc(node_size) for _ in groups]
        nxt = []
        for gi, (g, (addr, view)) in enumerate(zip(groups, addrs)):
            left = addrs[gi - 1][0] if gi > 0 else UNDEF
            right = addrs[gi + 1][0] if gi + 1 < len(groups) else UNDEF
            b = b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(g), left, right)
            for k, child, _ in g:
                b += k + struct.pack("<Q", child)
            end = groups[gi + 1][0][2] if gi + 1 < len(groups) else tuple(o + 1 for o in last_offs)
            b += key(0, end)                                  # closing key: first chunk past this node
            view[: len(b)] = b
            nxt.append((g[0][0], addr, g[0][2]))
        if len(nxt) == 1:
            return nxt[0][1]
        items, level = nxt, level + 1


def save(path: str, arrays: Dict[str, np.ndarray], chunks: Optional[Dict[str, Tuple[int, ...]]] = None,
         gzip: Optional[int] = None, shuffle: bool = False) -> None:
    """Write ``{name: array}`` as little-endian datasets of the root group (superblock version 0, one symbol node: at
    most 32 datasets — HDF5_OUTPUT writes two).  Datasets named in ``chunks`` are stored chunked (v1 chunk B-tree),
    optionally through the shuffle and deflate filters; the others contiguously."""
    names = sorted(arrays)                                   # symbol nodes are searched by name: sorted order
    if not names or len(names) > 32:
        raise HDF5Error("between 1 and 32 datasets per file")
    arrs = {}
    for n in names:
        a = np.ascontiguousarray(arrays[n])
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        if a.dtype == np.float16 or a.dtype.kind not in "fiu":
            a = a.astype(np.float32)
        arrs[n] = a
    leaf_k = 16
    O = 8
    sb_size = 24 + 4 * O + (2 * O + 24)                      # superblock v0 incl. the root symbol-table entry = 96
    # layout plan: superblock | root object header | local heap header | heap data | B-tree node | symbol node |
    #              dataset object headers | raw data
    root_hdr = sb_size
    root_msgs = _msg(0x11, struct.pack("<QQ", 0, 0))         # patched below
    root_hdr_size = 16 + len(root_msgs)
    heap = root_hdr + root_hdr_size
    heap_hdr_size = 8 + 2 * 8 + O
    heap_data = heap + heap_hdr_size
    blob = b"\0" * 8                                         # offset 0: the empty name (B-tree key 0)
    name_off = {}
    for n in names:
        name_off[n] = len(blob)
        blob += _pad8(n.encode() + b"\0")
    free_off = len(blob)
    blob += struct.pack("<QQ", 1, 16)                        # one free block: next = 1 (last), size 16
    btree = heap_data + len(blob)
    btree_size = 8 + 2 * O + (2 * leaf_k + 1) * 8 + 2 * leaf_k * O
    snod = btree + btree_size
    snod_size = 8 + 2 * leaf_k * (2 * O + 24)
    p = snod + snod_size
    chunks = chunks or {}

    def ds_msgs(n, a, addr):
        # the message set libhdf5 writes for a plain dataset: fill value (v1: late / incremental allocation, written if
        # set, default value), datatype, simple dataspace, [filter pipeline,] contiguous or chunked layout
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        if n not in chunks:
            fill = struct.pack("<BBBBI", 1, 2, 2, 1, 0)
            return _msg(0x05, fill, 1) + _msg(0x03, _dtype_msg(a.dtype), 1) + _msg(0x01, space) + \
                _msg(0x08, struct.pack("<BBQQ", 3, 1, addr, a.nbytes))
        fill = struct.pack("<BBBBI", 1, 3, 2, 1, 0)
        pipe = b""
        flt = ([(2, b"shuffle\0", (a.itemsize,))] if shuffle else []) + ([(1, b"deflate\0", (int(gzip),))] if gzip else [])
        if flt:
            body = struct.pack("<BB6x", 1, len(flt))
            for fid, fname, cd in flt:
                body += struct.pack("<HHHH", fid, len(fname), 1, len(cd)) + _pad8(fname)
                body += b"".join(struct.pack("<I", v) for v in cd) + (b"\0" * 4 if len(cd) % 2 else b"")
            pipe = _msg(0x0B, body, 1)
        cd = tuple(chunks[n]) + (a.itemsize,)
        layout = struct.pack("<BBBQ", 3, 2, len(cd), addr) + b"".join(struct.pack("<I", c) for c in cd)
        return _msg(0x05, fill, 1) + _msg(0x03, _dtype_msg(a.dtype), 1) + _msg(0x01, space) + pipe + _msg(0x08, layout)

    hdr_at = {}
    for n in names:
        if n in chunks and (len(chunks[n]) != arrs[n].ndim or arrs[n].ndim == 0 or min(chunks[n]) < 1):
            raise HDF5Error(f"dataset '{n}': chunk shape {chunks[n]} does not fit an array of shape {arrs[n].shape}")
        hdr_at[n] = p
        p += 16 + len(ds_msgs(n, arrs[n], 0))
    p = (p + 7) // 8 * 8
    data_at = {}
    tail = bytearray()                                       # everything after the headers, addressed from ``p``
    tail_at = p

    def put(raw: bytes) -> int:
        addr = tail_at + len(tail)
        tail.extend(raw + b"\0" * (-len(raw) % 8))
        return addr

    def alloc(nbytes):                                       # for the B-tree builder: (address, writable window)
        addr = put(b"\0" * nbytes)
        return addr, _Window(tail, addr - tail_at)

    for n in names:
        a = arrs[n]
        if n not in chunks:
            data_at[n] = put(a.tobytes())
            continue
        cs = tuple(chunks[n])
        grid = [range(0, d, c) for d, c in zip(a.shape, cs)]
        entries = []
        for idx in np.ndindex(*[len(g) for g in grid]):
            o = tuple(g[i] for g, i in zip(grid, idx))
            block = np.zeros(cs, a.dtype)                    # edge chunks are stored whole, padded with the fill value
            src = a[tuple(slice(x, x + c) for x, c in zip(o, cs))]
            block[tuple(slice(0, k) for k in src.shape)] = src
            raw = block.tobytes()
            if shuffle:
                raw = _shuffle(raw, a.itemsize)
            if gzip:
                raw = zlib.compress(raw, int(gzip))
            entries.append((len(raw), o, put(raw)))
        data_at[n] = _chunk_tree(entries, a.ndim, alloc)
    eof = tail_at + len(tail)
    out = bytearray(tail_at) + tail
    # superblock
    out[0:8] = SIGNATURE
    out[8:16] = bytes([0, 0, 0, 0, 0, O, 8, 0])
    out[16:24] = struct.pack("<HHI", leaf_k, 16, 0)
    out[24:56] = struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out[56:96] = struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap)
    # root object header
    root_msgs = _msg(0x11, struct.pack("<QQ", btree, heap))
    out[root_hdr: root_hdr + 16] = struct.pack("<BxHII4x", 1, 1, 1, len(root_msgs))
    out[root_hdr + 16: root_hdr + 16 + len(root_msgs)] = root_msgs
    # local heap
    out[heap: heap + heap_hdr_size] = b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(blob), free_off, heap_data)
    out[heap_data: heap_data + len(blob)] = blob
    # B-tree: one leaf (the symbol node); keys: 0 (empty name) and the largest name
    bt = b"TREE" + bytes([0, 0]) + struct.pack("<H", 1) + struct.pack("<QQ", UNDEF, UNDEF)
    bt += struct.pack("<QQQ", 0, snod, name_off[names[-1]])
    out[btree: btree + len(bt)] = bt
    # symbol node
    sn = b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(names))
    for n in names:
        sn += struct.pack("<QQII16x", name_off[n], hdr_at[n], 0, 0)
    out[snod: snod + len(sn)] = sn
    # datasets
    for n in names:
        msgs = ds_msgs(n, arrs[n], data_at[n])
        h = hdr_at[n]
        out[h: h + 16] = struct.pack("<BxHII4x", 1, _count_msgs(msgs), 1, len(msgs))
        out[h + 16: h + 16 + len(msgs)] = msgs
    with open(path, "wb") as fh:
        fh.write(out)
