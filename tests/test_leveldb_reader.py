"""Read-only LevelDB reader (C++ host runtime) vs a writer that follows leveldb's log / table / manifest format documents;
the snappy decoder is cross-checked against pyarrow's independent implementation."""
import os
import struct

import numpy as np
import pytest

from poseidon_b200 import proto as P
from poseidon_b200.data import native

pytestmark = pytest.mark.skipif(not native.available(), reason="host extension could not be built")


def varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def lp(b):
    return varint(len(b)) + b


def ikey(user_key, seq, typ=1):
    return user_key + struct.pack("<Q", (seq << 8) | typ)


def build_block(entries, restart_interval=4):
    """entries: sorted [(key, value)] -> prefix-compressed block with restart array."""
    out = bytearray()
    restarts = []
    prev = b""
    for i, (k, v) in enumerate(entries):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, entries, per_block=5, compress=False):
    """entries: sorted [(internal key, value)]."""
    import pyarrow as pa
    f = bytearray()
    index = []
    for i in range(0, len(entries), per_block):
        blk = build_block(entries[i:i + per_block])
        ctype = 0
        if compress:
            blk = pa.Codec("snappy").compress(blk, asbytes=True)
            ctype = 1
        off = len(f)
        f += blk + bytes([ctype]) + b"\0\0\0\0"                       # trailer: type + crc (not checked by the reader)
        index.append((entries[min(i + per_block, len(entries)) - 1][0], varint(off) + varint(len(blk))))
    meta = build_block([])
    moff = len(f)
    f += meta + b"\0" + b"\0\0\0\0"
    iblk = build_block(index, restart_interval=1)
    ioff = len(f)
    f += iblk + b"\0" + b"\0\0\0\0"
    footer = varint(moff) + varint(len(meta)) + varint(ioff) + varint(len(iblk))
    footer += b"\0" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57)
    f += footer
    with open(path, "wb") as fh:
        fh.write(f)
    return len(f)


def write_log(path, records):
    """leveldb log format: 32 KiB blocks, 7-byte headers, FULL / FIRST / MIDDLE / LAST fragments."""
    out = bytearray()
    for rec in records:
        pos, first = 0, True
        while True:
            left = 32768 - (len(out) % 32768)
            if left < 7:
                out += b"\0" * left
                continue
            n = min(len(rec) - pos, left - 7)
            last = pos + n == len(rec)
            typ = 1 if (first and last) else 2 if first else 4 if last else 3
            out += struct.pack("<IHB", 0, n, typ) + rec[pos:pos + n]
            pos += n
            first = False
            if last:
                break
    with open(path, "wb") as fh:
        fh.write(out)


def _datum(i, rng, shape=(3, 6, 6)):
    c, h, w = shape
    d = P.Datum(channels=c, height=h, width=w, label=i % 10)
    d.data = rng.randint(0, 256, c * h * w).astype(np.uint8).tobytes()
    return d.SerializeToString()


def make_db(path, compress):
    os.makedirs(path)
    rng = np.random.RandomState(11)
    truth = {}
    # table 5 (older): keys 0..39 at sequences 1..40
    t5 = []
    for i in range(40):
        k, v = f"{i:08d}".encode(), _datum(i, rng)
        t5.append((ikey(k, i + 1), v))
        truth[k] = v
    s5 = write_table(os.path.join(path, "000005.ldb"), t5, compress=compress)
    # table 7 (newer, .sst extension): overwrites keys 10..14, deletes key 3, adds keys 40..49 (one large value)
    t7 = []
    seq = 100
    for i in list(range(10, 15)) + list(range(40, 50)):
        k = f"{i:08d}".encode()
        v = _datum(i + 1000, rng, shape=(3, 64, 64) if i == 45 else (3, 6, 6))
        t7.append((ikey(k, seq), v))
        truth[k] = v
        seq += 1
    t7.append((ikey(b"00000003", seq, typ=0), b""))
    del truth[b"00000003"]
    t7.sort(key=lambda kv: (kv[0][:-8], -struct.unpack("<Q", kv[0][-8:])[0]))
    s7 = write_table(os.path.join(path, "000007.sst"), t7, compress=compress)
    # a dead table that the manifest deleted again
    write_table(os.path.join(path, "000004.ldb"), [(ikey(b"zzzz", 1), b"stale")])
    # write-ahead log 9: put key 50, overwrite key 0, delete key 49
    batch = struct.pack("<QI", 500, 3)
    v50, v0 = _datum(50, rng), _datum(7777, rng)
    batch += b"\x01" + lp(b"00000050") + lp(v50) + b"\x01" + lp(b"00000000") + lp(v0) + b"\x00" + lp(b"00000049")
    truth[b"00000050"], truth[b"00000000"] = v50, v0
    del truth[b"00000049"]
    big = struct.pack("<QI", 600, 1) + b"\x01" + lp(b"00000051") + lp(_datum(51, rng, shape=(3, 120, 120)))   # spans log blocks
    write_log(os.path.join(path, "000009.log"), [batch, big])
    write_log(os.path.join(path, "000002.log"), [struct.pack("<QI", 1, 1) + b"\x01" + lp(b"old") + lp(b"flushed long ago")])
    # MANIFEST: comparator, log number 9, new files 4 / 5 / 7, then delete 4
    e1 = varint(1) + lp(b"leveldb.BytewiseComparator") + varint(2) + varint(9) + varint(3) + varint(10) + varint(4) + varint(700)
    e1 += varint(7) + varint(1) + varint(4) + varint(100) + lp(ikey(b"zzzz", 1)) + lp(ikey(b"zzzz", 1))
    e1 += varint(7) + varint(1) + varint(5) + varint(s5) + lp(t5[0][0]) + lp(t5[-1][0])
    e2 = varint(7) + varint(0) + varint(7) + varint(s7) + lp(t7[0][0]) + lp(t7[-1][0]) + varint(6) + varint(1) + varint(4)
    e2 += varint(5) + varint(1) + lp(ikey(b"00000020", 5))
    write_log(os.path.join(path, "MANIFEST-000008"), [e1, e2])
    with open(os.path.join(path, "CURRENT"), "w") as fh:
        fh.write("MANIFEST-000008\n")
    # the value of key 51 as written into the batch
    p = 12 + 1
    klen = 8
    p += 1 + klen
    # decode the varint length of the value
    shift, vlen = 0, 0
    while True:
        b = big[p]
        p += 1
        vlen |= (b & 0x7f) << shift
        shift += 7
        if not b & 0x80:
            break
    truth[b"00000051"] = big[p:p + vlen]
    return truth


@pytest.mark.parametrize("compress", [False, True])
def test_leveldb_merge_of_tables_and_wal(tmp_path, compress):
    truth = make_db(str(tmp_path / "db"), compress)
    db = native.NativeRecordDB(str(tmp_path / "db"))
    keys = [db.key(i) for i in range(len(db))]
    assert keys == sorted(truth)                      # bytewise order, deleted keys gone, dead table ignored
    for i, k in enumerate(keys):
        assert db.value(i) == truth[k], k
    assert db.datum(keys.index(b"00000045")).height == 64


def test_snappy_decoder_matches_pyarrow():
    import pyarrow as pa
    rng = np.random.RandomState(5)
    m = native.module()
    for blob in (b"", b"a", b"abcabcabcabc" * 500, rng.randint(0, 4, 70000).astype(np.uint8).tobytes(),
                 rng.randint(0, 256, 3000).astype(np.uint8).tobytes(), bytes(200000)):
        comp = pa.Codec("snappy").compress(blob, asbytes=True)
        assert m.snappy_uncompress(comp) == blob
    with pytest.raises(RuntimeError):
        m.snappy_uncompress(b"\x05\xff\xff")


def test_data_layer_reads_leveldb(tmp_path):
    from poseidon_b200.net.net import Net
    truth = make_db(str(tmp_path / "train_leveldb"), False)
    f = tmp_path / "net.prototxt"
    f.write_text(f'''layers {{ name: "data" type: DATA top: "data" top: "label"
        data_param {{ source: "{tmp_path / "train_leveldb"}" backend: LEVELDB batch_size: 4 }} }}''')
    net = Net(P.read_net(str(f)), phase=P.TRAIN)
    _, outs = net.forward()
    ks = sorted(k for k in truth if P.Datum.FromString(truth[k]).height == 6)[:4]
    assert net.blobs["data"].shape[1:] == (3, 6, 6)
    assert outs["label"].reshape(-1).tolist() == [float(P.Datum.FromString(truth[k]).label) for k in sorted(truth)[:4]]
    net.close()
