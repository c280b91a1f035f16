"""Read-only LMDB parser vs a writer that follows the LMDB 0.9 on-disk definition (no liblmdb in the image)."""
import os

import numpy as np
import pytest

from poseidon_b200 import proto as P
from poseidon_b200.data.lmdb_reader import LMDBFile, LMDBFormatError

from poseidon_b200.data.lmdb_writer import write_lmdb  # noqa: E402


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


def test_convert_imageset_exports_lmdb(tmp_path):
    import cv2
    from poseidon_b200.tools import convert_imageset
    rng = np.random.RandomState(4)
    root = tmp_path / "img"
    os.makedirs(root)
    lines = []
    for i in range(6):
        cv2.imwrite(str(root / f"{i}.png"), rng.randint(0, 256, (9, 7, 3)).astype(np.uint8))
        lines.append(f"{i}.png {i % 3}")
    (tmp_path / "list.txt").write_text("\n".join(lines) + "\n")
    assert convert_imageset.main([str(root) + "/", str(tmp_path / "list.txt"), str(tmp_path / "out_lmdb"), "--backend", "lmdb"]) == 0
    db = LMDBFile(str(tmp_path / "out_lmdb"))
    assert len(db) == 6 and db.key(2).startswith(b"00000002_")
    d = db.datum(4)
    assert (d.channels, d.height, d.width, d.label) == (3, 9, 7, 1)
    want = cv2.imread(str(root / "4.png")).transpose(2, 0, 1)
    assert np.array_equal(np.frombuffer(d.data, np.uint8).reshape(3, 9, 7), want)
    db.close()
