"""LMDB (data.mdb) writer — the export counterpart of :mod:`lmdb_reader`.

Builds a fresh single-transaction environment bottom-up (leaf pages in key order, overflow runs for values larger than
half a page, branch levels on top, two meta pages) following the LMDB 0.9 on-disk definition.  Records must be supplied
in ascending bytewise key order (what ``convert_imageset`` produces with its zero-padded keys).  Round-trips with the
built-in reader; there is no liblmdb in the build image to cross-check against.
"""
import os
import struct

PSIZE = 4096
HDR = 16


def _page_hdr(pgno, flags, lower=0, upper=0, pages=None):
    if pages is not None:
        return struct.pack("<QHHI", pgno, 0, flags, pages)
    return struct.pack("<QHHHH", pgno, 0, flags, lower, upper)


def _meta(pgno, txnid, root, depth, entries, last_pg, branch, leaf, overflow):
    free_db = struct.pack("<IHHQQQQQ", PSIZE, 0, 0, 0, 0, 0, 0, 0xFFFFFFFFFFFFFFFF)
    main_db = struct.pack("<IHHQQQQQ", 0, 0, depth, branch, leaf, overflow, entries, root)
    body = struct.pack("<IIQQ", 0xBEEFC0DE, 1, 0, 1 << 30) + free_db + main_db + struct.pack("<QQ", last_pg, txnid)
    page = _page_hdr(pgno, 0x08) + body
    return page + b"\0" * (PSIZE - len(page))


def write_lmdb(path, records, max_leaf_nodes=None):
    """records: sorted list of (key bytes, value bytes).  Layout: meta0, meta1, leaves / overflow runs in key order, branch
    levels, root last.  Writes ``path/data.mdb`` (+ an empty ``lock.mdb``, which liblmdb recreates anyway)."""
    records = list(records)
    for (k0, _), (k1, _) in zip(records, records[1:]):
        if not k0 < k1:
            raise ValueError("LMDB writer: keys must be strictly ascending")
    os.makedirs(path, exist_ok=True)
    pages = {}                                  # pgno -> bytes
    next_pg = [2]

    def alloc(n=1):
        p = next_pg[0]
        next_pg[0] += n
        return p

    leaves = []                                 # (first key, pgno)
    overflow_pages = 0
    cur = []                                    # nodes of the leaf being filled: (key, node bytes)

    def flush_leaf():
        if not cur:
            return
        pg = alloc()
        body = bytearray(PSIZE)
        upper = PSIZE
        ptrs = []
        for _, node in cur:
            upper -= len(node) + (len(node) & 1)
            body[upper:upper + len(node)] = node
            ptrs.append(upper)
        lower = HDR + 2 * len(ptrs)
        body[:HDR] = _page_hdr(pg, 0x02, lower, upper)
        body[HDR:lower] = struct.pack(f"<{len(ptrs)}H", *ptrs)
        pages[pg] = bytes(body)
        leaves.append((cur[0][0], pg))
        cur.clear()

    used = 0
    for key, val in records:
        big = len(val) + len(key) + 8 > PSIZE // 2 - HDR
        if big:
            npg = (HDR + len(val) + PSIZE - 1) // PSIZE
            opg = alloc(npg)
            blob = _page_hdr(opg, 0x04, pages=npg) + val
            blob += b"\0" * (npg * PSIZE - len(blob))
            for i in range(npg):
                pages[opg + i] = blob[i * PSIZE:(i + 1) * PSIZE]
            overflow_pages += npg
            node = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, 0x01, len(key)) + key + struct.pack("<Q", opg)
        else:
            node = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, 0, len(key)) + key + val
        need = len(node) + (len(node) & 1) + 2
        if cur and (used + need > PSIZE - HDR or (max_leaf_nodes and len(cur) >= max_leaf_nodes)):
            flush_leaf()
            used = 0
        cur.append((key, node))
        used += need
    flush_leaf()

    depth, branch_pages = 1, 0
    level = leaves
    while len(level) > 1:
        nxt = []
        for i in range(0, len(level), 32):
            grp = level[i:i + 32]
            pg = alloc()
            body = bytearray(PSIZE)
            upper = PSIZE
            ptrs = []
            for j, (k, child) in enumerate(grp):
                kk = b"" if j == 0 else k           # the first branch key is implicit
                node = struct.pack("<HHHH", child & 0xFFFF, (child >> 16) & 0xFFFF, child >> 32, len(kk)) + kk
                upper -= len(node) + (len(node) & 1)
                body[upper:upper + len(node)] = node
                ptrs.append(upper)
            lower = HDR + 2 * len(ptrs)
            body[:HDR] = _page_hdr(pg, 0x01, lower, upper)
            body[HDR:lower] = struct.pack(f"<{len(ptrs)}H", *ptrs)
            pages[pg] = bytes(body)
            nxt.append((grp[0][0], pg))
            branch_pages += 1
        level = nxt
        depth += 1
    root = level[0][1] if level else 0xFFFFFFFFFFFFFFFF
    last = next_pg[0] - 1
    with open(os.path.join(path, "data.mdb"), "wb") as f:
        f.write(_meta(0, 1, root, depth if level else 0, len(records), last, branch_pages, len(leaves), overflow_pages))
        f.write(_meta(1, 0, 0xFFFFFFFFFFFFFFFF, 0, 0, 1, 0, 0, 0))          # older (empty) transaction
        for pg in range(2, next_pg[0]):
            f.write(pages[pg])
    open(os.path.join(path, "lock.mdb"), "wb").close()
