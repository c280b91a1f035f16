"""convert_imageset: "<image path> <label>" list -> record DB of serialized Datum protos.

    python -m poseidon_b200.tools.convert_imageset [--gray] [--shuffle] [--resize_height H --resize_width W]
        ROOTFOLDER/ LISTFILE DB_NAME

Output is the framework's PDB record store or (``--backend lmdb``) an LMDB environment; the Datum
payload is byte-identical to what the reference writes, keys are "%08d_<path>" as there).
reference: tools/convert_imageset.cpp:33-43 (flags), :60-160 (main loop), src/caffe/util/io.cpp:83-130.
"""
from __future__ import annotations

import argparse
import random
import sys

from .. import proto as P
from ..data.db import RecordWriter
from ..data.images import read_image


def image_to_datum(path, label, new_h=0, new_w=0, color=True):
    img = read_image(path, new_h, new_w, color)
    c, h, w = img.shape
    return P.Datum(channels=c, height=h, width=w, data=img.tobytes(), label=int(label))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("listfile")
    ap.add_argument("db")
    ap.add_argument("--gray", action="store_true")
    ap.add_argument("--shuffle", action="store_true")
    ap.add_argument("--backend", default="pdb")
    ap.add_argument("--resize_width", type=int, default=0)
    ap.add_argument("--resize_height", type=int, default=0)
    args = ap.parse_args(argv)
    lines = []
    with open(args.listfile) as f:
        for line in f:
            line = line.strip()
            if line:
                name, label = line.rsplit(None, 1)
                lines.append((name, int(label)))
    if args.shuffle:
        random.shuffle(lines)
    n = 0
    backend = args.backend.lower()
    if backend not in ("pdb", "lmdb", "leveldb"):
        raise SystemExit("--backend must be pdb (native record store), lmdb or leveldb (exports that stock Caffe / "
                         "Poseidon read; leveldb is the reference's default)")

    def records():
        nonlocal n
        for i, (name, label) in enumerate(lines):
            try:
                d = image_to_datum(args.root + name, label, args.resize_height, args.resize_width, not args.gray)
            except IOError as e:
                print(e, file=sys.stderr)
                continue
            n += 1
            if n % 1000 == 0:
                print(f"Processed {n} files.", file=sys.stderr)
            yield ("%08d_%s" % (i, name)).encode(), d.SerializeToString()

    if backend == "lmdb":
        # an LMDB environment (data.mdb) that the reference's DataLayer reads with `backend: LMDB`
        from ..data.lmdb_writer import write_lmdb
        write_lmdb(args.db, records())
    elif backend == "leveldb":
        from ..data.leveldb_writer import write_leveldb
        write_leveldb(args.db, records())
    else:
        with RecordWriter(args.db) as w:
            for key, value in records():
                w.put(key, value)
    print(f"Processed {n} files.", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
