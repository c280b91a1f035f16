"""data/leveldb_writer.py: bulk-load writer for LevelDB directories, read back by the C++ reader; checksums verified
independently in Python (masked CRC-32C of every block and log record)."""
import os
import struct

import numpy as np
import pytest

from poseidon_b200 import proto as P
from poseidon_b200.data import native
from poseidon_b200.data.db import open_db
from poseidon_b200.data.leveldb_writer import write_leveldb

pytestmark = pytest.mark.skipif(not native.available(), reason="host extension could not be built")


def _records(n, rng, big_every=0):
    out = []
    for i in range(n):
        shape = (3, 40, 40) if big_every and i % big_every == 0 else (3, 6, 6)
        d = P.Datum(channels=shape[0], height=shape[1], width=shape[2], label=i % 10)
        d.data = rng.randint(0, 256, int(np.prod(shape))).astype(np.uint8).tobytes()
        out.append((b"%08d" % i, d.SerializeToString()))
    return out


def _unmask(c):
    c = (c - 0xA282EAD8) & 0xFFFFFFFF
    return ((c >> 17) | (c << 15)) & 0xFFFFFFFF


def test_roundtrip_multi_table_and_checksums(tmp_path):
    rng = np.random.RandomState(0)
    recs = _records(700, rng, big_every=50)
    db = str(tmp_path / "db")
    assert write_leveldb(db, recs, table_bytes=40000, block_size=1024) == 700
    names = sorted(os.listdir(db))
    tables = [n for n in names if n.endswith(".ldb")]
    assert len(tables) > 3 and "CURRENT" in names and "MANIFEST-000004" in names and "LOCK" in names
    assert open(os.path.join(db, "CURRENT")).read() == "MANIFEST-000004\n"
    r = native.module().RecordDB(db)
    assert r.size() == 700
    for i in (0, 1, 49, 50, 333, 699):
        assert r.key(i) == recs[i][0] and r.value(i) == recs[i][1]
    # every table: footer magic, and the CRC of the index block + first data block verify
    crc = native.module().crc32c
    for t in tables:
        raw = open(os.path.join(db, t), "rb").read()
        assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57

        def varint(p):
            v = s = 0
            while True:
                b = raw[p]
                p += 1
                v |= (b & 0x7F) << s
                s += 7
                if not b & 0x80:
                    return v, p
        p = len(raw) - 48
        _, p = varint(p)
        _, p = varint(p)
        ioff, p = varint(p)
        ilen, p = varint(p)
        for off, ln in ((ioff, ilen),):
            stored = struct.unpack("<I", raw[off + ln + 1: off + ln + 5])[0]
            assert raw[off + ln] == 0 and _unmask(stored) == crc(raw[off: off + ln + 1])
    # MANIFEST: one FULL record whose CRC covers type + payload
    man = open(os.path.join(db, "MANIFEST-000004"), "rb").read()
    stored, ln, typ = struct.unpack("<IHB", man[:7])
    assert typ in (1, 2) and _unmask(stored) == crc(bytes([typ]) + man[7: 7 + ln])
    assert b"leveldb.BytewiseComparator" in man


def test_rejects_unsorted_and_overwrites_old_database(tmp_path):
    db = str(tmp_path / "db")
    with pytest.raises(ValueError, match="ascending"):
        write_leveldb(db, [(b"b", b"1"), (b"a", b"2")])
    write_leveldb(db, [(b"a", b"1"), (b"b", b"2"), (b"c", b"3")])
    write_leveldb(db, [(b"x", b"9")])                                     # a second bulk load replaces the first
    r = native.module().RecordDB(db)
    assert r.size() == 1 and r.key(0) == b"x" and r.value(0) == b"9"


def test_tools_write_leveldb_like_the_reference(tmp_path):
    """convert_imageset --backend leveldb -> DATA layer with backend: LEVELDB; extract_features --backend leveldb."""
    import cv2
    from poseidon_b200 import Net
    from poseidon_b200.proto import parse_text
    from poseidon_b200.tools import convert_imageset
    rng = np.random.RandomState(1)
    lines = []
    for i in range(6):
        cv2.imwrite(str(tmp_path / f"im{i}.png"), rng.randint(0, 256, size=(10, 12, 3)).astype(np.uint8))
        lines.append(f"im{i}.png {i % 3}")
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    db = str(tmp_path / "imgs_leveldb")
    assert convert_imageset.main([str(tmp_path) + "/", str(tmp_path / "list.txt"), db, "--resize_height", "8",
                                  "--resize_width", "8", "--backend", "leveldb"]) == 0
    assert os.path.exists(os.path.join(db, "CURRENT"))
    r = open_db(db, "LEVELDB")
    assert len(r) == 6 and r.datum(0).height == 8 and r.datum(4).label == 1
    txt = f'''layers {{ name: "d" type: DATA top: "data" top: "label"
              data_param {{ source: "{db}" batch_size: 3 backend: LEVELDB }} }}'''
    net = Net(parse_text(txt, P.NetParameter), phase=P.TRAIN)
    _, out = net.forward()
    assert out["data"].shape == (3, 3, 8, 8) and out["label"].reshape(-1).tolist() == [0, 1, 2]
    net.close()


@pytest.mark.parametrize("backend", ["leveldb", "lmdb", "pdb"])
def test_partition_data_writes_the_requested_backend(tmp_path, backend):
    """tools.partition_data: LevelDB in, N round-robin shards out in the reference's formats (it only supports LevelDB)."""
    from poseidon_b200.tools import partition_data
    rng = np.random.RandomState(2)
    recs = _records(10, rng)
    src = str(tmp_path / "src_leveldb")
    write_leveldb(src, recs)
    assert partition_data.main([src, "--num_partitions", "3", "--backend", backend]) == 0
    total = []
    for k in range(3):
        r = open_db(f"{src}_{k}", {"leveldb": "LEVELDB", "lmdb": "LMDB", "pdb": "LEVELDB"}[backend])
        assert len(r) == (4, 3, 3)[k]
        total += [(r.key(i), r.value(i)) for i in range(len(r))]
    assert sorted(total) == recs
