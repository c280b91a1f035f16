"""partition_data: round-robin split of a record DB into N shards ``<db>_0 .. <db>_{N-1}`` — the layout the
DATA layer expects when ``shared_file_system`` is false (each client opens ``source_<client_id>``).

    python -m poseidon_b200.tools.partition_data --num_partitions N DB_PATH
reference: tools/partition_data.cpp:30 (flags), :87-112 (round-robin copy), :120.
"""
from __future__ import annotations

import argparse
import sys

from ..data.db import RecordWriter, open_db


def partition(db_path: str, n: int, backend: str = "pdb"):
    """Input: any readable database (record store, LMDB, LevelDB).  Output shards: the record store, or — like the
    reference — LevelDB directories (``backend="leveldb"``) / LMDB environments (``"lmdb"``)."""
    db = open_db(db_path)
    backend = backend.lower()
    if backend == "leveldb":
        from ..data.leveldb_writer import LevelDBWriter as Writer
    elif backend == "lmdb":
        from ..data.lmdb_writer import write_lmdb

        class Writer:                                  # LMDB is bulk-loaded from sorted records, like LevelDB
            def __init__(self, path):
                self.path, self.records = path, {}

            def put(self, key, value):
                self.records[bytes(key)] = bytes(value)

            def close(self):
                write_lmdb(self.path, sorted(self.records.items()))
    elif backend == "pdb":
        Writer = RecordWriter
    else:
        raise SystemExit("--backend must be pdb, lmdb or leveldb")
    writers = [Writer(f"{db_path}_{k}") for k in range(n)]
    for i in range(len(db)):
        writers[i % n].put(db.key(i), db.value(i))
    for w in writers:
        w.close()
    return [w.path for w in writers]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("db")
    ap.add_argument("--num_partitions", type=int, required=True)
    ap.add_argument("--backend", default="pdb")
    args = ap.parse_args(argv)
    for p in partition(args.db, args.num_partitions, args.backend):
        print(p)
    return 0


if __name__ == "__main__":
    sys.exit(main())
