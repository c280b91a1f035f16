"""Solver semantics on CPU: lr policies, update rules through the solver, test nets, snapshots, outputs."""
import os

import pytest
import torch

from poseidon_b200 import get_solver
from poseidon_b200 import proto as P
from poseidon_b200.models import zoo
from poseidon_b200.solver.lr_policy import learning_rate
from smallnet import feed, make_data, small_net, small_solver_param


def test_lr_policies():
    sp = P.SolverParameter(base_lr=0.1, gamma=0.5, power=2.0, stepsize=10, max_iter=100)
    sp.lr_policy = "fixed"
    assert learning_rate(sp, 50) == pytest.approx(0.1)
    sp.lr_policy = "step"
    assert learning_rate(sp, 25) == pytest.approx(0.1 * 0.25)
    sp.lr_policy = "exp"
    assert learning_rate(sp, 3) == pytest.approx(0.1 * 0.125)
    sp.lr_policy = "inv"
    assert learning_rate(sp, 4) == pytest.approx(0.1 * (1 + 2.0) ** -2)
    sp.lr_policy = "poly"
    assert learning_rate(sp, 50) == pytest.approx(0.1 * 0.25)
    sp.lr_policy = "bogus"
    with pytest.raises(ValueError):
        learning_rate(sp, 1)


def _cpu_solver(solver_type="SGD", momentum=0.9, steps=3, **kw):
    net = small_net(batch=8, hw=19, with_lrn=True)
    sp = small_solver_param(net, max_iter=steps, solver_type=solver_type, momentum=momentum)
    sp.solver_mode = "CPU"
    for k, v in kw.items():
        setattr(sp, k, v)
    s = get_solver(sp)
    x, y = make_data(8 * 8, hw=19)
    feed(s, x, y)
    return s


@pytest.mark.parametrize("solver_type,momentum", [("SGD", 0.9), ("NESTEROV", 0.9), ("ADAGRAD", 0.0)])
def test_solver_step_matches_manual_update(solver_type, momentum):
    s = _cpu_solver(solver_type, momentum)
    w0 = [p.detach().clone() for p in s.net.params]
    # manual gradient on the same batch
    s.net.layer_by_name["data"].pos = 0
    loss, _ = s.net.forward()
    loss.backward()
    grads = [p.grad.detach().clone() for p in s.net.params]
    s.net.zero_grad_()
    s.net.layer_by_name["data"].pos = 0
    s.step(1)
    lr, wd = 0.01, 0.0005
    for p, w, g, lrm, wdm in zip(s.net.params, w0, grads, s.net.params_lr, s.net.params_weight_decay):
        gg = g + wd * wdm * w
        if solver_type == "SGD":
            step = lr * lrm * gg
        elif solver_type == "NESTEROV":
            step = (1 + momentum) * lr * lrm * gg
        else:
            step = lr * lrm * gg / (gg.abs() + 1e-8)
        assert torch.allclose(p.detach(), w - step, atol=1e-6), solver_type
    assert s.net.params_lr[1] == 2.0 and s.net.params_weight_decay[1] == 0.0      # bias multipliers from the prototxt


def test_adagrad_rejects_momentum():
    with pytest.raises(ValueError, match="Momentum cannot be used with AdaGrad"):
        _cpu_solver("ADAGRAD", 0.9)


def test_training_reduces_loss_and_test_net_scores(caplog):
    net = zoo.lenet(batch=16, test_batch=16)
    sp = zoo.get_solver_param("lenet", net=net, max_iter=40, display=10, test_interval=20, solver_mode="CPU",
                              snapshot=0, snapshot_after_train=False, random_seed=2, base_lr=0.02)
    sp.test_iter = [2]
    s = get_solver(sp)
    import logging
    with caplog.at_level(logging.INFO, logger="poseidon_b200"):
        s.solve()
    rows = s.train_table.rows
    assert rows[0][2] > rows[-1][2] > 0                      # loss column went down on the 4-batch synthetic pool
    assert any("Test net output #0: accuracy" in r.message for r in caplog.records)
    assert any(r.message.startswith("Iteration 0, loss:") for r in caplog.records)
    assert len(s.test_tables[0].rows) == 3                   # iters 0, 20, 40


def test_net_outputs_csv(tmp_path):
    s = _cpu_solver(steps=4, display=2)
    s.solve()
    out = tmp_path / "o.netoutputs"
    s.print_net_outputs(str(out))
    lines = out.read_text().strip().splitlines()
    assert lines[0] == "Iteration,time,loss,loss,"
    assert lines[1].startswith("0,") and lines[2].startswith("2,")


def test_snapshot_restore_resumes_exactly(tmp_path):
    def make():
        s = _cpu_solver(steps=6)
        s.param.snapshot_prefix = str(tmp_path / "snap")
        return s
    a = make()
    a.step(3)
    a.snapshot()
    a.step(3)
    ref = [p.detach().clone() for p in a.net.params]
    b = make()
    b.restore(str(tmp_path / "snap_iter_3.solverstate"))
    assert b.iter == 3
    b.net.layer_by_name["data"].pos = 3 * 8
    b.step(3)
    for p, q in zip(ref, b.net.params):
        assert torch.allclose(p, q.detach(), atol=1e-7)
    st = P.read_binary(str(tmp_path / "snap_iter_3.solverstate"), P.SolverState)
    assert st.iter == 3 and st.learned_net.endswith("snap_iter_3.caffemodel") and len(st.history) == len(a.net.params)


def test_finetune_weights(tmp_path):
    a = _cpu_solver(steps=2)
    a.step(2)
    a.param.snapshot_prefix = str(tmp_path / "ft")
    model, _ = a.snapshot()
    b = _cpu_solver(steps=2)
    b.load_weights(model)
    for p, q in zip(a.net.params, b.net.params):
        assert torch.equal(p.detach(), q.detach())
    assert b.iter == 0


def test_solver_requires_a_net():
    with pytest.raises(ValueError, match="must specify a train net"):
        get_solver(P.SolverParameter(base_lr=0.1, lr_policy="fixed", max_iter=1))


def test_caffe_main_cli_train_and_time(tmp_path):
    from poseidon_b200.tools import caffe_main
    net_path = tmp_path / "lenet.prototxt"
    P.write_text(str(net_path), zoo.lenet(batch=8, test_batch=8))
    sp = zoo.lenet_solver(net_path=str(net_path), max_iter=4, display=2, test_interval=4, solver_mode="CPU", snapshot=4,
                          snapshot_prefix=str(tmp_path / "lenet"))
    sp.test_iter = [1]
    solver_path = tmp_path / "solver.prototxt"
    P.write_text(str(solver_path), sp)
    rc = caffe_main.main(["train", f"--solver={solver_path}", "--svb=true", "--table_staleness=0",
                          "--num_comm_channels_per_client=8", "--consistency_model=SSPPush",
                          f"--net_outputs={tmp_path / 'out'}"])
    assert rc == 0
    assert os.path.exists(tmp_path / "lenet_iter_4.caffemodel") and os.path.exists(tmp_path / "out.netoutputs")
    rc = caffe_main.main(["test", f"--model={net_path}", f"--weights={tmp_path / 'lenet_iter_4.caffemodel'}",
                          "--iterations=2", "--gpu=-1"])
    assert rc == 0
    import logging
    records = []

    class Grab(logging.Handler):
        def emit(self, r):
            records.append(r.getMessage())
    h = Grab()
    lg = logging.getLogger("poseidon_b200")
    old_level = lg.level
    lg.addHandler(h)
    lg.setLevel(logging.INFO)
    try:
        assert caffe_main.main(["time", f"--model={net_path}", "--iterations=1", "--gpu=-1"]) == 0
    finally:
        lg.removeHandler(h)
        lg.setLevel(old_level)
    # per-layer forward AND backward lines, like the reference's `caffe time` (tools/caffe_main.cpp:255-328)
    fw = [m for m in records if " forward: " in m]
    bw = [m for m in records if " backward: " in m]
    assert len(fw) == len(bw) > 5 and bw[0].startswith("loss") and fw[0].startswith("data")
    assert float([m for m in bw if m.startswith("conv2")][0].split("backward:")[1].split()[0]) > 0
    assert caffe_main.main(["device_query"]) == 0


def test_jsonl_metrics(tmp_path, monkeypatch):
    import json
    path = tmp_path / "m" / "metrics.jsonl"
    monkeypatch.setenv("POSEIDON_METRICS_JSONL", str(path))
    net = zoo.lenet(batch=4, test_batch=4)
    sp = zoo.get_solver_param("lenet", net=net, max_iter=6, display=2, solver_mode="CPU", snapshot=0,
                              snapshot_after_train=False, test_interval=0, random_seed=5)
    sp.clear("test_iter")
    s = get_solver(sp, engine="torch")
    s.solve()
    rows = [json.loads(l) for l in open(path)]
    assert [r["iter"] for r in rows] == [0, 2, 4]
    assert "images_per_sec_per_rank" not in rows[0] and rows[1]["images_per_sec_per_rank"] > 0
    assert rows[2]["world_size"] == 1 and "loss" in rows[2]["outputs"]
    s.close()


def test_foreach_deferred_step_matches_per_tensor_rule():
    """The CUDA-graph-safe optimizer of the library backends (device-resident lr, torch._foreach kernels, one step per
    iteration) is Caffe's SGD rule exactly: same weights / history as gradsync.apply_rule, incl. lr / decay multipliers."""
    import torch
    from poseidon_b200.parallel import gradsync as G
    torch.manual_seed(3)
    hy = G.Hyper(G.SGD, momentum=0.9, weight_decay=5e-4)
    hy.lr = 0.01
    ws = [torch.randn(7, 5), torch.randn(5), torch.randn(3, 2, 2, 2)]
    gs = [torch.randn_like(w) for w in ws]
    hs = [torch.rand_like(w) * 0.1 for w in ws]
    mults = [(1.0, 1.0), (2.0, 0.0), (1.0, 1.0)]
    ref_w, ref_h = [w.clone() for w in ws], [h.clone() for h in hs]
    for w, g, h, (lm, dm) in zip(ref_w, gs, ref_h, mults):
        G.apply_rule(hy, w, g.clone(), h, lm, dm, decay_scale=2.0)
    params = [torch.nn.Parameter(w.clone()) for w in ws]
    hist = [h.clone() for h in hs]
    lr_t = torch.tensor(0.01)
    G.foreach_sgd_step(hy, [(p, g.clone(), h, lm, dm) for p, g, h, (lm, dm) in zip(params, gs, hist, mults)], lr_t,
                       decay_scale=2.0)
    for p, h, rw, rh in zip(params, hist, ref_w, ref_h):
        assert torch.allclose(p.data, rw, atol=1e-6) and torch.allclose(h, rh, atol=1e-6)
