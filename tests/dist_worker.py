"""One rank of a small data-parallel training job (launched by test_dist_*.py with RANK/WORLD_SIZE env).
Writes rank 0's (and optionally every rank's) final weights to --out."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch

from poseidon_b200 import get_solver, init_rank_context
from smallnet import feed, make_data, small_net, small_solver_param


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--engine", default="torch")
    ap.add_argument("--comm", default="auto")
    ap.add_argument("--svb", type=int, default=0)
    ap.add_argument("--sfb_mode", default="all")
    ap.add_argument("--staleness", type=int, default=0)
    ap.add_argument("--grad_reduce", default="sum")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=19)
    ap.add_argument("--base_lr", type=float, default=0.01)
    ap.add_argument("--device", default=None)
    ap.add_argument("--delay_rank", type=int, default=-1)
    ap.add_argument("--solver_type", default="SGD")
    ap.add_argument("--aggr_fraction", type=float, default=0.1)
    ap.add_argument("--wire_dtype", default=None)
    ap.add_argument("--freeze", default="", help="comma separated layers whose blobs get blobs_lr 0 (finetuning)")
    ap.add_argument("--snapshot_prefix", default="", help="write <prefix>_iter_N.{caffemodel,solverstate} after the last step")
    ap.add_argument("--restore", default="", help="resume from this .solverstate; --steps counts the steps still to run")
    ap.add_argument("--total_steps", type=int, default=0, help="with --restore: length of the uninterrupted run (data order)")
    ap.add_argument("--net", default="small", choices=["small", "inception"], help="inception: branchy net (lanes)")
    ap.add_argument("--graph", type=int, default=0, help="capture the training step as a CUDA graph")
    args = ap.parse_args()
    rc = init_rank_context(args.device)
    M, W = args.batch, rc.world_size
    classes = 16
    if args.net == "inception":
        from test_lanes import inception_net
        args.hw, classes = 20, 10
        net = inception_net(batch=M, classes=classes, hw=args.hw)
    else:
        net = small_net(batch=M, hw=args.hw)
    for l in net.layers:
        if l.name in args.freeze.split(","):
            l.blobs_lr = [0.0, 0.0]
    sp = small_solver_param(net, base_lr=args.base_lr, max_iter=args.steps, solver_type=args.solver_type,
                            momentum=0.0 if args.solver_type == "ADAGRAD" else 0.9)
    if rc.device.type == "cpu":
        sp.solver_mode = "CPU"
    s = get_solver(sp, rank_ctx=rc, engine=args.engine, comm=args.comm, svb=bool(args.svb), sfb_mode=args.sfb_mode,
                   staleness=args.staleness, grad_reduce=args.grad_reduce, aggr_fraction=args.aggr_fraction,
                   wire_dtype=args.wire_dtype,
                   dtype=torch.float32 if args.engine == "torch" else None)
    total = args.total_steps or args.steps
    first = total - args.steps if args.restore else 0
    x, y = make_data(M * W * total, classes=classes, hw=args.hw)
    # global batch t = samples [t*M*W, (t+1)*M*W); this rank takes the slice [r*M, (r+1)*M) of it
    idx = torch.cat([torch.arange(t * M * W + rc.rank * M, t * M * W + (rc.rank + 1) * M)
                     for t in range(first, first + args.steps)])
    feed(s, x[idx], y[idx])
    if args.snapshot_prefix:
        sp.snapshot_prefix = args.snapshot_prefix
    if args.restore:
        s.restore(args.restore)
    if args.delay_rank == rc.rank and hasattr(s.sync.backend, "delay_hook"):
        import time
        s.sync.backend.delay_hook = lambda clock: time.sleep(0.05)
    if args.graph:
        s.enable_cuda_graph(warmup=1)               # 1 eager + 1 captured step
        s.step(args.steps - 2)
    else:
        s.step(args.steps)
    s.sync.wait_all()
    if hasattr(s.sync.backend, "drain"):
        s.sync.backend.drain()
    if rc.device.type == "cuda":
        torch.cuda.synchronize()
    rc.barrier()
    if args.snapshot_prefix:
        s.snapshot()
        rc.barrier()
    w = {f"{n}.{j}": l.export_blob(j) for n, l in zip(s.net.layer_names, s.net.layers) for j in range(len(l.blobs))}
    extra = {}
    if hasattr(s.sync.backend, "max_observed_lag"):
        extra["max_lag"] = np.array(s.sync.backend.max_observed_lag)
    wire = s.sync.backend.bytes_on_wire()
    for k, v in wire.items():
        extra["wire_" + k] = np.array(v)
    np.savez(f"{args.out}.{rc.rank}.npz", loss=np.array(float(s.last_loss)), **w, **extra)
    s.close()
    rc.shutdown()


if __name__ == "__main__":
    main()
