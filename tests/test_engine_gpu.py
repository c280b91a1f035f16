"""Whole-net checks of the sm100 engine against the fp32 torch engine on identical data/weights."""
import pytest
import torch

from smallnet import feed, make_data, small_net, small_solver_param

pytestmark = pytest.mark.gpu


def _run(engine, steps, solver_type="SGD", momentum=0.9):
    from poseidon_b200 import get_solver
    net = small_net(batch=16)
    sp = small_solver_param(net, max_iter=steps, solver_type=solver_type, momentum=momentum)
    s = get_solver(sp, engine=engine, dtype=torch.float32 if engine == "torch" else None)
    x, y = make_data(16 * steps)
    feed(s, x, y)
    losses = []
    for _ in range(steps):
        s.step(1)
        losses.append(float(s.last_loss))
    torch.cuda.synchronize()
    weights = {f"{n}.{j}": l.export_blob(j) for n, l in zip(s.net.layer_names, s.net.layers)
               for j in range(len(l.blobs))}
    s.close()
    return losses, weights


@pytest.mark.parametrize("solver_type,momentum", [("SGD", 0.9), ("NESTEROV", 0.9), ("ADAGRAD", 0.0)])
def test_sm100_matches_fp32_engine(ext, solver_type, momentum):
    steps = 4
    l_ref, w_ref = _run("torch", steps, solver_type, momentum)
    l_sm, w_sm = _run("sm100", steps, solver_type, momentum)
    for a, b in zip(l_ref, l_sm):
        assert abs(a - b) < 0.05 * max(1.0, abs(a)), (l_ref, l_sm)
    import numpy as np
    if solver_type == "ADAGRAD":
        return      # the first AdaGrad steps are ±lr·sign(g): weight-level comparison is dominated by bf16 sign flips
    for name in w_ref:
        d = np.abs(w_ref[name] - w_sm[name]).max()
        m = np.abs(w_ref[name]).max()
        # biases start at a constant: compare the *change* they underwent, not their absolute value
        init = 0.1 if name in ("conv1.1", "conv2.1", "fc4.1") else 0.0
        mm = np.abs(w_ref[name] - init).max() if name.endswith(".1") else m
        assert d <= 0.08 * mm + 3e-4, f"{name}: max diff {d} vs magnitude {mm}"


def test_smoke_entry(ext):
    import __graft_entry__ as g
    g.smoke()


def test_snapshot_roundtrip_sm100(ext, tmp_path):
    from poseidon_b200 import get_solver
    net = small_net(batch=16)
    sp = small_solver_param(net, max_iter=2)
    sp.snapshot_prefix = str(tmp_path / "small")
    s = get_solver(sp, engine="sm100")
    x, y = make_data(64)
    feed(s, x, y)
    s.step(2)
    s.snapshot()
    ref = {n: l.export_blob(0) for n, l in zip(s.net.layer_names, s.net.layers) if len(l.blobs)}
    s2 = get_solver(sp, engine="torch", dtype=torch.float32)
    s2.restore(str(tmp_path / "small_iter_2.solverstate"))
    import numpy as np
    for n, l in zip(s2.net.layer_names, s2.net.layers):
        if len(l.blobs):
            assert np.allclose(l.export_blob(0), ref[n], atol=1e-6), n
    assert s2.iter == 2


def test_cuda_graph_step_matches_eager(ext):
    """The captured step (forward+backward+fused updates as one CUDA graph) follows the eager trajectory."""
    from poseidon_b200 import get_solver

    def run(graph):
        net = small_net(batch=16)
        sp = small_solver_param(net, max_iter=8)
        s = get_solver(sp, engine="sm100")
        x, y = make_data(16 * 8)
        feed(s, x, y)
        if graph:
            s.enable_cuda_graph(warmup=2)          # 2 eager + 1 captured step
            s.step(5)
        else:
            s.step(8)
        torch.cuda.synchronize()
        out = {n: l.export_blob(0) for n, l in zip(s.net.layer_names, s.net.layers) if len(l.blobs)}
        loss = float(s.last_loss)
        s.close()
        return out, loss

    import numpy as np
    w_e, l_e = run(False)
    w_g, l_g = run(True)
    assert abs(l_e - l_g) < 0.05 * max(1.0, abs(l_e)), (l_e, l_g)
    for n in w_e:
        d = np.abs(w_e[n] - w_g[n]).max()
        assert d <= 0.02 * np.abs(w_e[n]).max() + 1e-4, (n, d)


def test_googlenet_small_batch_step(ext):
    """GoogLeNet (59 conv, 9 concat, 3 losses) trains on the sm100 engine."""
    from poseidon_b200 import get_solver
    from poseidon_b200.models import zoo
    net = zoo.googlenet(batch=4, test_batch=4)
    sp = zoo.get_solver_param("googlenet", net=net, display=0, snapshot=0, snapshot_after_train=False, test_interval=0,
                              max_iter=2, random_seed=3)
    sp.clear("test_iter")
    sp.base_lr = 0.0005
    s = get_solver(sp, engine="sm100")
    s.step(1)
    first = float(s.last_loss)
    s.step(2)
    loss = float(s.last_loss)
    # (0.3 + 0.3 + 1) * ln(1000) = 11.05 with calm logits; dropout 0.7 in the aux heads raises it a little
    assert 8 < first < 22, first
    assert loss == loss and loss < 60, loss
    s.close()


def test_solve_replays_the_step_as_a_cuda_graph(ext, tmp_path):
    """``Solver.solve`` — the path behind ``caffe_main train`` and ``CaffeEngine.start`` — captures the step as a CUDA
    graph by itself (display / snapshot boundaries included) and ends where the eager run ends."""
    from poseidon_b200 import get_solver

    def run(graph, sub):
        net = small_net(batch=16)
        sp = small_solver_param(net, max_iter=12)
        sp.display = 5
        sp.snapshot = 10
        sp.snapshot_prefix = str(tmp_path / sub / "small")
        (tmp_path / sub).mkdir()
        s = get_solver(sp, engine="sm100")
        s.use_cuda_graph = graph
        x, y = make_data(16 * 12)
        feed(s, x, y)
        s.solve()
        torch.cuda.synchronize()
        out = {n: l.export_blob(0) for n, l in zip(s.net.layer_names, s.net.layers) if len(l.blobs)}
        graphed = getattr(s, "_graph", None) is not None
        rows = [r[0] for r in s.train_table.rows]
        s.close()
        return out, graphed, rows

    import numpy as np
    import os
    w_e, g_e, rows_e = run(False, "eager")
    w_g, g_g, rows_g = run(None, "graph")                 # None = auto
    assert not g_e and g_g
    assert rows_e == rows_g == [0, 5, 10]                  # every display iteration was displayed in both modes
    assert os.path.exists(tmp_path / "graph" / "small_iter_10.solverstate")
    for n in w_e:
        d = np.abs(w_e[n] - w_g[n]).max()
        assert d <= 0.02 * np.abs(w_e[n]).max() + 1e-4, (n, d)


def test_vendor_arm_cuda_graph_matches_eager(ext):
    """The constructed vendor baseline (torch engine: cuDNN/cuBLAS + foreach SGD with a device-side learning rate)
    can be replayed as one CUDA graph and follows its own eager trajectory."""
    from poseidon_b200 import get_solver

    def run(graph):
        net = small_net(batch=16)
        sp = small_solver_param(net, max_iter=8)
        s = get_solver(sp, engine="torch", dtype=torch.float32)
        x, y = make_data(16 * 8)
        feed(s, x, y)
        if graph:
            s.enable_cuda_graph(warmup=2)
            s.step(5)
        else:
            s.step(8)
        torch.cuda.synchronize()
        out = {n: l.export_blob(0) for n, l in zip(s.net.layer_names, s.net.layers) if len(l.blobs)}
        s.close()
        return out

    import numpy as np
    w_e, w_g = run(False), run(True)
    for n in w_e:
        assert np.allclose(w_e[n], w_g[n], atol=2e-4, rtol=1e-3), n
