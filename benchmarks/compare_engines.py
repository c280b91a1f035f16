"""Layer-by-layer forward comparison of the sm100 engine against the fp32 torch engine (same weights, same batch)."""
import sys
import torch
sys.path.insert(0, ".")
from poseidon_b200 import get_solver
from poseidon_b200.models import zoo

model = sys.argv[1] if len(sys.argv) > 1 else "googlenet"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)


def build(engine):
    net = zoo.get_model(model, batch=batch, test_batch=batch)
    sp = zoo.get_solver_param(model, net=net, display=0, snapshot=0, snapshot_after_train=False, test_interval=0,
                              max_iter=1, random_seed=3)
    sp.clear("test_iter")
    s = get_solver(sp, engine=engine, dtype=torch.float32 if engine == "torch" else None)
    for l in s.net.layers:            # deterministic: no dropout noise
        if l.type_name == "DROPOUT":
            l.ctx = type("C", (), {"train": False, "engine": l.ctx.engine, "device": l.ctx.device})()
    return s

a, b = build("torch"), build("sm100")
# identical weights: copy canonical blobs torch -> sm100
for la, lb in zip(a.net.layers, b.net.layers):
    for j in range(len(la.blobs)):
        lb.import_blob(j, la.export_blob(j))
    st = getattr(lb, "_sm100", None)
    if st is not None:
        st.mark_updated()
        if getattr(st, "arena_shadow", False) or st.wb is not None:
            st.dirty_wb = True
# identical input: run torch data layer, feed both
with torch.no_grad():
    data = a.net.forward_data()
    la, _ = a.net.forward(dict(data), start=a.net.num_leading_data_layers())
    lb, _ = b.net.forward({k: v.clone() for k, v in data.items()}, start=b.net.num_leading_data_layers())
print("loss torch", float(la), "sm100", float(lb))
worst = []
for name in a.net.blob_shapes:
    if name in a.net.blobs and name in b.net.blobs:
        x, y = a.net.blobs[name].float(), b.net.blobs[name].float()
        if x.shape != y.shape:
            print("shape mismatch", name, x.shape, y.shape)
            continue
        err = (x - y).abs().max().item()
        mag = x.abs().max().item() + 1e-6
        worst.append((err / mag, name, err, mag))
bad = [w for w in worst if w[0] > 0.05]
print("blobs compared", len(worst), "bad", len(bad))
for w in worst[:0] + bad[:15]:
    print("  %-40s rel %.3f  err %.4g  mag %.4g" % (w[1], w[0], w[2], w[3]))
