"""One optimizer step on both engines from identical weights/data: relative error of every blob's UPDATE."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from poseidon_b200 import get_solver
from smallnet import feed, make_data, small_net, small_solver_param

def run(engine, steps=1):
    net = small_net(batch=16)
    sp = small_solver_param(net, max_iter=steps, momentum=0.9)
    s = get_solver(sp, engine=engine, dtype=torch.float32 if engine == "torch" else None)
    x, y = make_data(16 * steps)
    feed(s, x, y)
    w0 = {f"{n}.{j}": l.export_blob(j).copy() for n, l in zip(s.net.layer_names, s.net.layers) for j in range(len(l.blobs))}
    s.step(steps)
    torch.cuda.synchronize()
    w1 = {f"{n}.{j}": l.export_blob(j).copy() for n, l in zip(s.net.layer_names, s.net.layers) for j in range(len(l.blobs))}
    return w0, w1, float(s.last_loss)

for steps in (1, 3):
    a0, a1, la = run("torch", steps)
    b0, b1, lb = run("sm100", steps)
    print(f"steps={steps} loss torch {la:.5f} sm100 {lb:.5f}")
    for k in a0:
        da, db = a1[k] - a0[k], b1[k] - b0[k]
        init = np.abs(a0[k] - b0[k]).max()
        rel = np.linalg.norm(da - db) / (np.linalg.norm(da) + 1e-12)
        print(f"   {k:10s} init diff {init:.2e}  |upd| {np.linalg.norm(da):.4e}  rel err of update {rel:.4f}  ratio {np.linalg.norm(db)/ (np.linalg.norm(da)+1e-12):.4f}")
