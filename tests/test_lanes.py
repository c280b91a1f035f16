"""Branch-parallel execution plan (net/lanes.py): static lane assignment on CPU, stream execution parity on the GPU."""
import numpy as np
import pytest
import torch

from poseidon_b200 import proto as P
from poseidon_b200.models.zoo import NetBuilder


def inception_net(batch=8, classes=10, hw=20, aux=True):
    """Two inception-style modules (4 branches + zero-copy CONCAT) and an auxiliary loss head; no dropout."""
    b = NetBuilder("lanesnet")
    b.layer("data", "MEMORY_DATA", (), ("data", "label"),
            memory_data_param={"batch_size": batch, "channels": 3, "height": hw, "width": hw})
    g = {"type": "gaussian", "std": 0.05}
    c0 = {"type": "constant", "value": 0.1}

    def cr(name, bottom, nout, k, pad=0):
        b.conv(name, bottom, nout, k, pad=pad, wf=g, bf=c0)
        b.relu(name + "/relu", name)
        return name

    x = cr("stem", "data", 32, 3, 1)

    def module(tag, bottom, c1, r3, c3, r5, c5, pp):
        p = f"inc{tag}/"
        cr(p + "1x1", bottom, c1, 1)
        cr(p + "3x3r", bottom, r3, 1)
        cr(p + "3x3", p + "3x3r", c3, 3, 1)
        cr(p + "5x5r", bottom, r5, 1)
        cr(p + "5x5", p + "5x5r", c5, 5, 2)
        b.pool(p + "pool", bottom, "MAX", 3, 1, 1)
        cr(p + "proj", p + "pool", pp, 1)
        return b.layer(p + "out", "CONCAT", (p + "1x1", p + "3x3", p + "5x5", p + "proj"), (p + "out",))

    x = module("A", x, 16, 16, 32, 8, 16, 16)
    if aux:
        b.pool("aux/pool", x, "AVE", 5, 3)
        b.fc("aux/fc", "aux/pool", classes, wf=g, bf=c0)
        b.softmax_loss("aux/loss", "aux/fc", top="aux/loss", weight=0.3)
    x = module("B", x, 32, 16, 32, 8, 16, 16)
    b.pool("pool", x, "MAX", 2, 2)
    b.fc("fc", "pool", classes, wf=g, bf=c0)
    b.softmax_loss("loss", "fc")
    return b.net


def test_lane_plan_forks_the_inception_branches():
    from poseidon_b200.layers import NetContext
    from poseidon_b200.net.lanes import default_lanes, plan_lanes
    from poseidon_b200.net.net import Net
    ctx = NetContext(phase=P.TRAIN, device="cpu", engine="torch")
    assert default_lanes(ctx) == 1                      # the library arm keeps the sequential schedule
    net = Net(inception_net(), phase=P.TRAIN, ctx=ctx)
    assert net.n_lanes == 1 and net._lane_runner(0, len(net.layers) - 1) is None
    plan_lanes(net, 4)
    assert net.n_lanes == 4
    lane = {n: l for n, l in zip(net.layer_names, net.lane)}
    for tag in ("A", "B"):
        heads = [lane[f"inc{tag}/1x1"], lane[f"inc{tag}/3x3r"], lane[f"inc{tag}/5x5r"], lane[f"inc{tag}/pool"]]
        assert sorted(heads) == [0, 1, 2, 3], heads     # four branches, four streams
        assert lane[f"inc{tag}/3x3"] == lane[f"inc{tag}/3x3r"] and lane[f"inc{tag}/5x5"] == lane[f"inc{tag}/5x5r"]
        assert lane[f"inc{tag}/proj"] == lane[f"inc{tag}/pool"]
        i = net.layer_names.index(f"inc{tag}/out")
        assert lane[f"inc{tag}/out"] == heads[0]        # the CONCAT continues its first bottom's lane ...
        assert len(net.lane_wait[i]) == 3               # ... and waits for the three others
    # every cross-lane edge has its producer's top registered for the consumer's stream
    for i, waits in enumerate(net.lane_wait):
        for p in waits:
            assert any(net.lane[i] in share for share in net.lane_share[p]), (net.layer_names[p], net.layer_names[i])
    # a chain has nothing to fork
    from smallnet import small_net
    chain = Net(small_net(), phase=P.TRAIN, ctx=NetContext(phase=P.TRAIN, device="cpu", engine="torch"))
    plan_lanes(chain, 4)
    assert chain.n_lanes == 1


def _train(lanes, graph, monkeypatch, steps=6, wgrad_lane=1, defer=1):
    from poseidon_b200 import get_solver
    monkeypatch.setenv("POSEIDON_LANES", str(lanes))
    monkeypatch.setenv("POSEIDON_WGRAD_LANE", str(wgrad_lane))
    monkeypatch.setenv("POSEIDON_WGRAD_DEFER", str(defer))
    net = inception_net()
    sp = P.SolverParameter(base_lr=0.01, lr_policy="fixed", momentum=0.9, weight_decay=0.0005, display=0, max_iter=steps,
                           snapshot=0, snapshot_after_train=False, random_seed=11, solver_type="SGD", solver_mode="GPU")
    sp.net_param = net
    s = get_solver(sp, engine="sm100")
    assert s.net.n_lanes == (lanes if lanes > 1 else 1)
    rng = np.random.RandomState(5)
    x = torch.from_numpy(rng.randn(8 * steps, 3, 20, 20).astype(np.float32))
    y = torch.from_numpy(rng.randint(0, 10, size=(8 * steps,)).astype(np.float32))
    for layer in s.net.layers:
        if layer.type_name == "MEMORY_DATA":
            layer.reset(x, y)
    losses = []
    if graph:
        s.enable_cuda_graph(warmup=2)
        assert s._graph is not None
        s.step(steps - 3)
        losses.append(float(s.last_loss))
    else:
        for _ in range(steps):
            s.step(1)
            losses.append(float(s.last_loss))
    torch.cuda.synchronize()
    w = {n: l.export_blob(0) for n, l in zip(s.net.layer_names, s.net.layers) if len(l.blobs)}
    s.close()
    return losses, w


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
def test_lanes_match_the_sequential_schedule(ext, monkeypatch, graph):
    """Same kernels, same order per stream: the branch-parallel step must reproduce the one-stream trajectory (weights
    after 6 momentum steps; the per-weight reductions are order-independent, so the match is tight)."""
    l1, w1 = _train(1, graph, monkeypatch, wgrad_lane=0)          # one stream for everything
    # branch lanes; + weight gradients on side streams joined per layer; + joined at the update launches (default)
    for lanes, wl, defer in ((4, 0, 0), (4, 1, 0), (4, 1, 1), (1, 1, 1)):
        l4, w4 = _train(lanes, graph, monkeypatch, wgrad_lane=wl, defer=defer)
        assert np.allclose(l1, l4, rtol=2e-3, atol=2e-3), (lanes, wl, defer, l1, l4)
        for n in w1:
            d = np.abs(w1[n] - w4[n]).max()
            assert d <= 2e-3 * np.abs(w1[n]).max() + 1e-5, (lanes, wl, defer, n, d)


def test_lane_plan_of_the_zoo_nets():
    """GoogLeNet (train phase): every inception module forks into four distinct lanes and its CONCAT waits for three
    producers; the auxiliary classifier heads leave the trunk's lane; AlexNet / VGG-16 are chains and stay on one stream."""
    from poseidon_b200.layers import NetContext
    from poseidon_b200.models import zoo
    from poseidon_b200.net.lanes import plan_lanes
    from poseidon_b200.net.net import Net

    def build(param):
        return Net(param, phase=P.TRAIN, ctx=NetContext(phase=P.TRAIN, device="cpu", engine="torch",
                                                        data_shape_hint=(3, 224, 224)))
    g = build(zoo.googlenet(batch=2, test_batch=2))
    plan_lanes(g, 8)
    assert g.n_lanes == 8
    lane = dict(zip(g.layer_names, g.lane))
    for tag in ("3a", "3b", "4a", "4b", "4c", "4d", "4e", "5a", "5b"):
        p = f"inception_{tag}/"
        heads = {lane[p + "1x1"], lane[p + "3x3_reduce"], lane[p + "5x5_reduce"], lane[p + "pool"]}
        assert len(heads) == 4, (tag, heads)
        assert len(g.lane_wait[g.layer_names.index(p + "output")]) == 3
    for idx, trunk in ((1, "inception_4b/1x1"), (2, "inception_4e/1x1")):
        assert lane[f"loss{idx}/ave_pool"] != lane[trunk]
    for name in ("alexnet", "vgg16"):
        n = build(getattr(zoo, name)(batch=2, test_batch=2))
        plan_lanes(n, 8)
        assert n.n_lanes == 1, name
