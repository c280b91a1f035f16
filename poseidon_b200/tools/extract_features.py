"""extract_features: run a TEST net over N mini-batches on every rank and dump the named blobs (+ label) as
``Datum`` records, one DB per (blob, rank) — the distributed feature extractor.

    torchrun ... -m poseidon_b200.tools.extract_features PRETRAINED.caffemodel NET.prototxt blob1[,blob2] \
        out_db1[,out_db2] NUM_MINI_BATCHES [--gpu 0]

Rank 0 loads the weights and broadcasts them (the reference loads into the PS on client0/thread0 and every
worker syncs from the tables).
reference: src/caffe/feature_extractor.cpp:16-135, tools/extract_features.cpp:47-119.
"""
from __future__ import annotations

import argparse
import os
import sys



def main(argv=None):
    import torch
    from .. import proto as P
    from ..data.db import RecordWriter
    from ..layers import NetContext
    from ..net.net import Net
    from ..parallel.context import init_rank_context
    ap = argparse.ArgumentParser()
    ap.add_argument("weights")
    ap.add_argument("model")
    ap.add_argument("blobs")
    ap.add_argument("dbs")
    ap.add_argument("num_mini_batches", type=int)
    ap.add_argument("--gpu", default="")
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--backend", default="pdb", choices=["pdb", "leveldb"],
                    help="feature databases: the framework's record store, or LevelDB directories like the reference writes")
    args = ap.parse_args(argv)
    rc = init_rank_context(None if args.gpu != "-1" else "cpu")
    engine = args.engine if args.engine != "auto" else ("sm100" if rc.device.type == "cuda" else "torch")
    ctx = NetContext(phase=P.TEST, device=rc.device, engine=engine,
                     dtype=torch.bfloat16 if engine == "sm100" else torch.float32, rank=rc.rank,
                     world_size=rc.world_size, model_dir=os.path.dirname(os.path.abspath(args.model)))
    net = Net(P.read_net(args.model), phase=P.TEST, ctx=ctx)
    net.to(rc.device)
    if rc.is_root:
        net.copy_trained_layers_from(args.weights)
    for p in net.params:
        rc.broadcast_(p.data, 0)
    for l in net.layers:
        if getattr(l, "_sm100", None) is not None:
            l._sm100.mark_updated()
    blob_names = args.blobs.split(",")
    db_names = args.dbs.split(",")
    if len(blob_names) != len(db_names):
        raise SystemExit("the number of blob names and db names must be equal")
    for b in blob_names:
        if b not in net.blob_shapes:
            raise SystemExit(f"Unknown feature blob name {b} in the network {args.model}")
    if args.backend == "leveldb":
        from ..data.leveldb_writer import LevelDBWriter as Writer
    else:
        Writer = RecordWriter
    writers = [Writer(f"{d}_{rc.rank}_0") for d in db_names]
    idx = [0] * len(blob_names)
    with torch.no_grad():
        for _ in range(args.num_mini_batches):
            net.forward()
            label = net.blobs.get("label")
            lab = label.float().reshape(-1).cpu().numpy() if label is not None else None
            for k, name in enumerate(blob_names):
                feat = net.blobs[name].float()
                feat = feat.contiguous().cpu().numpy()
                n = feat.shape[0]
                f4 = feat.reshape(n, *((feat.shape[1:] + (1, 1, 1))[:3]))
                for i in range(n):
                    d = P.Datum(channels=f4.shape[1], height=f4.shape[2], width=f4.shape[3])
                    d.float_data = f4[i].reshape(-1)
                    if lab is not None:
                        d.label = int(lab[i])
                    writers[k].put("%d" % idx[k], d.SerializeToString())
                    idx[k] += 1
    for w in writers:
        w.close()
    rc.barrier()
    if rc.is_root:
        print("Successfully extracted the features!", file=sys.stderr)
    rc.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
