"""A small AlexNet-shaped net (conv/relu/LRN/pool/grouped conv/IP/softmax-loss, no dropout) fed by
MEMORY_DATA, used to compare engines and communication backends on identical inputs."""
import numpy as np
import torch

from poseidon_b200 import proto as P
from poseidon_b200.models.zoo import NetBuilder


def small_net(batch=16, classes=16, hw=35, with_lrn=True):
    b = NetBuilder("smallnet")
    b.layer("data", "MEMORY_DATA", (), ("data", "label"),
            memory_data_param={"batch_size": batch, "channels": 3, "height": hw, "width": hw})
    g = {"type": "gaussian", "std": 0.05}
    b.conv("conv1", "data", 32, 5, stride=2, wf=g, bf={"type": "constant", "value": 0.1})
    b.relu("relu1", "conv1")
    x = "conv1"
    if with_lrn:
        x = b.lrn("norm1", x, 5, 1e-2, 0.75)
    b.pool("pool1", x, "MAX", 3, 2)
    b.conv("conv2", "pool1", 64, 3, pad=1, group=2, wf=g, bf={"type": "constant", "value": 0.1})
    b.relu("relu2", "conv2")
    b.conv("conv3", "conv2", 64, 3, pad=1, wf=g, bf={"type": "constant", "value": 0.0})
    b.relu("relu3", "conv3")
    b.pool("pool3", "conv3", "MAX", 2, 2)
    b.fc("fc4", "pool3", 128, wf=g, bf={"type": "constant", "value": 0.1})
    b.relu("relu4", "fc4")
    b.fc("fc5", "fc4", classes, wf=g, bf={"type": "constant", "value": 0.0})
    b.softmax_loss("loss", "fc5")
    return b.net


def small_solver_param(net, base_lr=0.01, max_iter=4, solver_type="SGD", momentum=0.9):
    sp = P.SolverParameter(base_lr=base_lr, lr_policy="fixed", momentum=momentum, weight_decay=0.0005, display=0,
                           max_iter=max_iter, snapshot=0, snapshot_after_train=False, random_seed=11,
                           solver_type=solver_type, solver_mode="GPU")
    sp.net_param = net
    return sp


def make_data(n, classes=16, hw=35, seed=5):
    rng = np.random.RandomState(seed)
    x = rng.randn(n, 3, hw, hw).astype(np.float32)
    y = rng.randint(0, classes, size=(n,)).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(y)


def feed(solver, x, y):
    for layer in solver.net.layers:
        if layer.type_name == "MEMORY_DATA":
            layer.reset(x, y)
