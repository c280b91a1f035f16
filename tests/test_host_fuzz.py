"""Corrupted inputs to the C++ host parsers (record store, LMDB, LevelDB tables / logs / MANIFEST, snappy, LibSVM): every
outcome must be either a successful open or a Python exception — never a crash.  Under ``scripts/sanitize_host.sh`` the
same mutations run against an AddressSanitizer / UBSan build of the module, which turns silent out-of-bounds reads into
failures."""
import os
import shutil

import numpy as np
import pytest

from poseidon_b200 import proto as P
from poseidon_b200.data import native
from poseidon_b200.data.db import RecordWriter
from poseidon_b200.data.lmdb_writer import write_lmdb
from test_leveldb_reader import make_db

N_MUT = int(os.environ.get("POSEIDON_FUZZ_ITERS", "60"))


def _mutations(raw: bytes, rng, n):
    for i in range(n):
        b = bytearray(raw)
        kind = i % 4
        if kind == 0:                                       # flip a few bytes
            for _ in range(rng.randint(1, 6)):
                b[rng.randint(0, len(b))] = rng.randint(0, 256)
        elif kind == 1:                                     # truncate
            b = b[: rng.randint(0, len(b))]
        elif kind == 2:                                     # overwrite a run with 0xff (huge varints / lengths)
            a = rng.randint(0, len(b))
            b[a: a + rng.randint(1, 12)] = b"\xff" * min(12, len(b) - a)
        else:                                               # zero a run
            a = rng.randint(0, len(b))
            b[a: a + rng.randint(1, 40)] = b"\0" * min(40, len(b) - a)
        yield bytes(b)


def _touch_all(db):
    n = db.size()
    for i in range(min(n, 64)):
        db.key(i)
        db.value(i)


def _datums(n, rng):
    out = []
    for i in range(n):
        d = P.Datum(channels=3, height=5, width=5, label=i % 7)
        d.data = rng.randint(0, 256, 75).astype(np.uint8).tobytes()
        out.append((b"%08d" % i, d.SerializeToString()))
    return out


def _survives(open_fn):
    try:
        _touch_all(open_fn())
    except (RuntimeError, IndexError, ValueError, OSError, MemoryError):
        pass


def test_fuzz_record_store(tmp_path):
    rng = np.random.RandomState(0)
    src = tmp_path / "ok.pdb"
    with RecordWriter(str(src)) as w:
        for k, v in _datums(30, rng):
            w.put(k, v)
    raw = src.read_bytes()
    m = native.module()
    for i, mut in enumerate(_mutations(raw, rng, N_MUT)):
        p = tmp_path / "m.pdb"
        p.write_bytes(mut)
        _survives(lambda: m.RecordDB(str(p)))
        try:
            ld = m.BatchLoader(str(p), 4, 0, 1, 2)
            ld.stop()
        except (RuntimeError, ValueError):
            pass


def test_fuzz_lmdb(tmp_path):
    rng = np.random.RandomState(1)
    write_lmdb(str(tmp_path / "ok"), _datums(200, rng), max_leaf_nodes=7)
    raw = (tmp_path / "ok" / "data.mdb").read_bytes()
    m = native.module()
    os.makedirs(tmp_path / "m")
    for mut in _mutations(raw, rng, N_MUT):
        (tmp_path / "m" / "data.mdb").write_bytes(mut)
        _survives(lambda: m.RecordDB(str(tmp_path / "m")))


@pytest.mark.parametrize("compress", [False, True])
def test_fuzz_leveldb(tmp_path, compress):
    rng = np.random.RandomState(2)
    make_db(str(tmp_path / "ok"), compress)
    files = sorted(os.listdir(tmp_path / "ok"))
    m = native.module()
    per_file = max(4, N_MUT // len(files))
    for name in files:
        raw = (tmp_path / "ok" / name).read_bytes()
        for mut in _mutations(raw, rng, per_file):
            d = tmp_path / "m"
            if d.exists():
                shutil.rmtree(d)
            shutil.copytree(tmp_path / "ok", d)
            (d / name).write_bytes(mut)
            _survives(lambda: m.RecordDB(str(d)))


def test_fuzz_snappy_and_libsvm():
    rng = np.random.RandomState(3)
    m = native.module()
    # a valid snappy stream: literal + copies (hand-assembled: "abcdabcdabcdabcd....")
    payload = b"abcd" * 40
    good = bytes([0xA0, 0x01]) + bytes([(4 - 1) << 2]) + b"abcd" + bytes([((60 - 1) << 2) | 2, 4, 0]) * 2 + bytes([((36 - 1) << 2) | 2, 4, 0])
    assert m.snappy_uncompress(good) == payload
    for mut in _mutations(good, rng, N_MUT * 2):
        try:
            m.snappy_uncompress(mut)
        except (RuntimeError, MemoryError):
            pass
    text = b"".join(b"%d 1:0.5 7:1e-3 20:-4\n" % (i % 2) for i in range(50))
    for mut in _mutations(text, rng, N_MUT * 2):
        try:
            lab, ptr, idx, val = m.parse_libsvm(mut)
            assert len(ptr) == len(lab) + 1 and ptr[-1] == len(idx) == len(val)
        except RuntimeError:
            pass
