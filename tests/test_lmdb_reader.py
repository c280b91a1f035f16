"""Read-only LMDB parser vs a writer that follows the LMDB 0.9 on-disk definition (no liblmdb in the image)."""
import os
import struct

import numpy as np
import pytest

from poseidon_b200 import proto as P
from poseidon_b200.data.lmdb_reader import LMDBFile, LMDBFormatError

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
    """records: sorted list of (key bytes, value bytes).  Layout: meta0, meta1, leaves..., overflow runs..., branch root."""
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


def _records(n, shape, rng):
    out = []
    c, h, w = shape
    for i in range(n):
        d = P.Datum(channels=c, height=h, width=w, label=int(rng.randint(0, 10)))
        d.data = rng.randint(0, 256, c * h * w).astype(np.uint8).tobytes()
        out.append((f"{i:08d}".encode(), d.SerializeToString()))
    return out


def test_small_values_multi_level_tree(tmp_path):
    rng = np.random.RandomState(0)
    recs = _records(300, (1, 4, 4), rng)                      # ~40-byte values: many per leaf, two branch levels at 4/leaf
    write_lmdb(str(tmp_path / "db"), recs, max_leaf_nodes=4)
    db = LMDBFile(str(tmp_path / "db"))
    assert len(db) == 300 and db.depth == 3
    assert [k for k, _ in db] == [k for k, _ in recs]
    assert db.value(123) == recs[123][1] and db.key(299) == b"00000299"
    d = db.datum(7)
    assert (d.channels, d.height, d.width) == (1, 4, 4)
    db.close()


def test_overflow_values_like_imagenet(tmp_path):
    rng = np.random.RandomState(1)
    recs = _records(9, (3, 40, 40), rng)                      # 4.8 KB payloads -> F_BIGDATA + 2-page overflow runs
    write_lmdb(str(tmp_path / "db"), recs)
    db = LMDBFile(str(tmp_path / "db"))
    assert len(db) == 9
    for i, (k, v) in enumerate(db):
        assert k == recs[i][0] and v == recs[i][1]
    db.close()


def test_open_db_and_data_layer_read_lmdb(tmp_path):
    from poseidon_b200.data.db import open_db
    from poseidon_b200.net.net import Net
    rng = np.random.RandomState(2)
    recs = _records(10, (3, 8, 8), rng)
    write_lmdb(str(tmp_path / "train_lmdb"), recs)
    r = open_db(str(tmp_path / "train_lmdb"), "LMDB")
    assert len(r) == 10 and r.datum(3).label == P.Datum.FromString(recs[3][1]).label
    f = tmp_path / "net.prototxt"
    f.write_text(f'''layers {{ name: "data" type: DATA top: "data" top: "label"
        data_param {{ source: "{tmp_path / "train_lmdb"}" backend: LMDB batch_size: 4 }} }}''')
    net = Net(P.read_net(str(f)), phase=P.TRAIN)
    _, outs = net.forward()
    want = [P.Datum.FromString(v).label for _, v in recs[:4]]
    assert outs["label"].reshape(-1).tolist() == want
    net.close()


def test_rejects_garbage(tmp_path):
    p = tmp_path / "db"
    os.makedirs(p)
    (p / "data.mdb").write_bytes(b"\0" * 8192)
    with pytest.raises(LMDBFormatError):
        LMDBFile(str(p))


def test_native_loader_reads_lmdb(tmp_path):
    """The C++ batch loader indexes data.mdb with the same parser and fills batches identical to the Python path."""
    import torch
    from poseidon_b200.data import native
    from poseidon_b200.data.source import DBSource
    if not native.available():
        pytest.skip("host extension could not be built")
    rng = np.random.RandomState(3)
    big = _records(11, (3, 40, 40), rng)                    # overflow pages
    write_lmdb(str(tmp_path / "big"), big)
    small = _records(50, (1, 4, 4), rng)
    write_lmdb(str(tmp_path / "small"), small, max_leaf_nodes=4)
    for name, batch in (("big", 4), ("small", 8)):
        py = DBSource(LMDBFile(str(tmp_path / name)), batch, 1, 2)
        nat = native.NativeDBSource(str(tmp_path / name), batch, 1, 2, threads=2, depth=3, pin=False)
        for _ in range(5):
            xa, ya = py.next_batch()
            xb, yb = nat.next_batch()
            assert torch.equal(xa, xb) and torch.equal(ya, yb)
        nat.close()
