"""Model zoo: programmatic builders that emit V1-syntax ``NetParameter`` / ``SolverParameter``
for the model families the reference ships (LeNet, CIFAR-10 quick/full, AlexNet, CaffeNet,
GoogLeNet) plus VGG-16 (SFB-crossover stress config, BASELINE.json config #5).

The reference ships these as hand-written prototxts (models/bvlc_alexnet/train_val.prototxt,
models/bvlc_reference_caffenet/train_val.prototxt, models/bvlc_googlenet/train_test.prototxt,
examples/mnist/lenet_train_test.prototxt, examples/cifar10/cifar10_{quick,full}_train_test.prototxt);
user prototxts in that syntax are consumed as-is by :mod:`poseidon_b200.net`.  Here the same
architectures are *generated* so the repo needs no dataset-path-laden text files and so
batch size / data source / class count are parameters.
"""
from __future__ import annotations


from .. import proto as P
from ..proto import to_text


class NetBuilder:
    """Tiny NetSpec: append layers, get a NetParameter."""

    def __init__(self, name: str):
        self.net = P.NetParameter(name=name)

    def layer(self, name, typ, bottoms=(), tops=(), phase=None, lr=None, decay=None, loss_weight=None,
              **params):
        lp = P.LayerParameter(name=name, type=typ)
        lp.bottom = list(bottoms)
        lp.top = list(tops)
        if phase is not None:
            lp.include.add(phase=phase)
        if lr is not None:
            lp.blobs_lr = lr
        if decay is not None:
            lp.weight_decay = decay
        if loss_weight is not None:
            lp.loss_weight = [loss_weight]
        for sub, fields in params.items():
            m = lp.mutable(sub)
            for k, v in fields.items():
                if isinstance(v, dict):
                    mm = m.mutable(k)
                    for kk, vv in v.items():
                        setattr(mm, kk, vv)
                else:
                    setattr(m, k, v)
        self.net.layers.append(lp)
        return tops[0] if tops else None

    # -- common building blocks ----------------------------------------------------------
    def data(self, source, batch, test_source=None, test_batch=None, crop=0, mirror=False, mean_values=None,
             mean_file=None, scale=None, backend="LMDB"):
        for phase, src, bs, mir in ((P.TRAIN, source, batch, mirror), (P.TEST, test_source or source,
                                                                       test_batch or batch, False)):
            tp = {}
            if crop:
                tp["crop_size"] = crop
            if mir:
                tp["mirror"] = True
            if mean_values:
                tp["mean_value"] = list(mean_values)
            if mean_file:
                tp["mean_file"] = mean_file
            if scale is not None:
                tp["scale"] = scale
            kw = {"data_param": {"source": src, "batch_size": bs, "backend": backend,
                                 "shared_file_system": True}}
            if tp:
                kw["transform_param"] = tp
            self.layer("data", "DATA", (), ("data", "label"), phase=phase, **kw)

    def conv(self, name, bottom, nout, k, stride=1, pad=0, group=1, wf=None, bf=None, top=None,
             lr=(1, 2), decay=(1, 0)):
        cp = {"num_output": nout, "kernel_size": k}
        if stride != 1:
            cp["stride"] = stride
        if pad:
            cp["pad"] = pad
        if group != 1:
            cp["group"] = group
        cp["weight_filler"] = wf or {"type": "xavier"}
        cp["bias_filler"] = bf or {"type": "constant", "value": 0.0}
        return self.layer(name, "CONVOLUTION", (bottom,), (top or name,), lr=lr, decay=decay,
                          convolution_param=cp)

    def relu(self, name, blob):
        return self.layer(name, "RELU", (blob,), (blob,))

    def pool(self, name, bottom, method, k, stride=1, pad=0):
        pp = {"pool": method, "kernel_size": k, "stride": stride}
        if pad:
            pp["pad"] = pad
        return self.layer(name, "POOLING", (bottom,), (name,), pooling_param=pp)

    def lrn(self, name, bottom, size=5, alpha=1e-4, beta=0.75, region=None):
        p = {"local_size": size, "alpha": alpha, "beta": beta}
        if region:
            p["norm_region"] = region
        return self.layer(name, "LRN", (bottom,), (name,), lrn_param=p)

    def fc(self, name, bottom, nout, wf=None, bf=None, lr=(1, 2), decay=(1, 0)):
        ip = {"num_output": nout, "weight_filler": wf or {"type": "xavier"},
              "bias_filler": bf or {"type": "constant", "value": 0.0}}
        return self.layer(name, "INNER_PRODUCT", (bottom,), (name,), lr=lr, decay=decay,
                          inner_product_param=ip)

    def dropout(self, name, blob, ratio):
        return self.layer(name, "DROPOUT", (blob,), (blob,), dropout_param={"dropout_ratio": ratio})

    def softmax_loss(self, name, bottom, top=None, weight=None):
        return self.layer(name, "SOFTMAX_LOSS", (bottom, "label"), (top or name,), loss_weight=weight)

    def accuracy(self, name, bottom, top_k=1, phase=P.TEST):
        kw = {"accuracy_param": {"top_k": top_k}} if top_k != 1 else {}
        return self.layer(name, "ACCURACY", (bottom, "label"), (name,), phase=phase, **kw)


def _g(std):
    return {"type": "gaussian", "std": std}


def _c(v):
    return {"type": "constant", "value": v}


# ---------------------------------------------------------------------------------------------
def lenet(batch=64, test_batch=100, source="examples/mnist/mnist_train_lmdb",
          test_source="examples/mnist/mnist_test_lmdb"):
    """reference: examples/mnist/lenet_train_test.prototxt (20-50 conv, 500-10 ip, scale 1/256)."""
    b = NetBuilder("LeNet")
    b.data(source, batch, test_source, test_batch, scale=0.00390625)
    b.conv("conv1", "data", 20, 5)
    b.pool("pool1", "conv1", "MAX", 2, 2)
    b.conv("conv2", "pool1", 50, 5)
    b.pool("pool2", "conv2", "MAX", 2, 2)
    b.fc("ip1", "pool2", 500)
    b.relu("relu1", "ip1")
    b.fc("ip2", "ip1", 10)
    b.accuracy("accuracy", "ip2")
    b.softmax_loss("loss", "ip2")
    return b.net


def lenet_solver(net_path="lenet_train_test.prototxt", **over):
    """reference: examples/mnist/lenet_solver.prototxt."""
    sp = P.SolverParameter(net=net_path, base_lr=0.01, momentum=0.9, weight_decay=0.0005, lr_policy="inv",
                           gamma=0.0001, power=0.75, display=100, max_iter=10000, snapshot=5000,
                           snapshot_prefix="lenet", solver_mode="GPU", test_interval=500)
    sp.test_iter = [100]
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


def cifar10_quick(batch=100, source="examples/cifar10/cifar10_train_leveldb",
                  test_source="examples/cifar10/cifar10_test_leveldb"):
    """reference: examples/cifar10/cifar10_quick_train_test.prototxt."""
    b = NetBuilder("CIFAR10_quick")
    b.data(source, batch, test_source, batch, mean_file="mean.binaryproto", backend="LEVELDB")
    b.conv("conv1", "data", 32, 5, pad=2, wf=_g(0.0001), bf=_c(0), decay=None)
    b.pool("pool1", "conv1", "MAX", 3, 2)
    b.relu("relu1", "pool1")
    b.conv("conv2", "pool1", 32, 5, pad=2, wf=_g(0.01), bf=_c(0), decay=None)
    b.relu("relu2", "conv2")
    b.pool("pool2", "conv2", "AVE", 3, 2)
    b.conv("conv3", "pool2", 64, 5, pad=2, wf=_g(0.01), bf=_c(0), decay=None)
    b.relu("relu3", "conv3")
    b.pool("pool3", "conv3", "AVE", 3, 2)
    b.fc("ip1", "pool3", 64, wf=_g(0.1), bf=_c(0), decay=None)
    b.fc("ip2", "ip1", 10, wf=_g(0.1), bf=_c(0), decay=None)
    b.accuracy("accuracy", "ip2")
    b.softmax_loss("loss", "ip2")
    return b.net


def cifar10_full(batch=100, source="examples/cifar10/cifar10_train_leveldb",
                 test_source="examples/cifar10/cifar10_test_leveldb"):
    """reference: examples/cifar10/cifar10_full_train_test.prototxt (within-channel LRN)."""
    b = NetBuilder("CIFAR10_full")
    b.data(source, batch, test_source, batch, mean_file="mean.binaryproto", backend="LEVELDB")
    b.conv("conv1", "data", 32, 5, pad=2, wf=_g(0.0001), bf=_c(0), decay=None)
    b.pool("pool1", "conv1", "MAX", 3, 2)
    b.relu("relu1", "pool1")
    b.lrn("norm1", "pool1", 3, 5e-05, 0.75, "WITHIN_CHANNEL")
    b.conv("conv2", "norm1", 32, 5, pad=2, wf=_g(0.01), bf=_c(0), decay=None)
    b.relu("relu2", "conv2")
    b.pool("pool2", "conv2", "AVE", 3, 2)
    b.lrn("norm2", "pool2", 3, 5e-05, 0.75, "WITHIN_CHANNEL")
    b.conv("conv3", "norm2", 64, 5, pad=2, wf=_g(0.01), bf=_c(0), decay=None)
    b.relu("relu3", "conv3")
    b.pool("pool3", "conv3", "AVE", 3, 2)
    b.fc("ip1", "pool3", 10, wf=_g(0.01), bf=_c(0), decay=(250, 0))
    b.accuracy("accuracy", "ip1")
    b.softmax_loss("loss", "ip1")
    return b.net


def _alex_like(name, batch, test_batch, classes, norm_before_pool: bool, source, test_source):
    b = NetBuilder(name)
    b.data(source, batch, test_source, test_batch, crop=227, mirror=True,
           mean_file="data/ilsvrc12/imagenet_mean.binaryproto")
    b.conv("conv1", "data", 96, 11, stride=4, wf=_g(0.01), bf=_c(0))
    b.relu("relu1", "conv1")
    if norm_before_pool:       # AlexNet: conv -> relu -> norm -> pool
        b.lrn("norm1", "conv1")
        b.pool("pool1", "norm1", "MAX", 3, 2)
        x = "pool1"
    else:                      # CaffeNet: conv -> relu -> pool -> norm
        b.pool("pool1", "conv1", "MAX", 3, 2)
        b.lrn("norm1", "pool1")
        x = "norm1"
    b.conv("conv2", x, 256, 5, pad=2, group=2, wf=_g(0.01), bf=_c(0.1 if norm_before_pool else 1))
    b.relu("relu2", "conv2")
    if norm_before_pool:
        b.lrn("norm2", "conv2")
        b.pool("pool2", "norm2", "MAX", 3, 2)
        x = "pool2"
    else:
        b.pool("pool2", "conv2", "MAX", 3, 2)
        b.lrn("norm2", "pool2")
        x = "norm2"
    hi = 0.1 if norm_before_pool else 1
    b.conv("conv3", x, 384, 3, pad=1, wf=_g(0.01), bf=_c(0))
    b.relu("relu3", "conv3")
    b.conv("conv4", "conv3", 384, 3, pad=1, group=2, wf=_g(0.01), bf=_c(hi))
    b.relu("relu4", "conv4")
    b.conv("conv5", "conv4", 256, 3, pad=1, group=2, wf=_g(0.01), bf=_c(hi))
    b.relu("relu5", "conv5")
    b.pool("pool5", "conv5", "MAX", 3, 2)
    b.fc("fc6", "pool5", 4096, wf=_g(0.005), bf=_c(hi))
    b.relu("relu6", "fc6")
    b.dropout("drop6", "fc6", 0.5)
    b.fc("fc7", "fc6", 4096, wf=_g(0.005), bf=_c(hi))
    b.relu("relu7", "fc7")
    b.dropout("drop7", "fc7", 0.5)
    b.fc("fc8", "fc7", classes, wf=_g(0.01), bf=_c(0))
    b.accuracy("accuracy", "fc8")
    b.softmax_loss("loss", "fc8")
    return b.net


def alexnet(batch=256, test_batch=50, classes=1000, source="ilsvrc12_train_lmdb",
            test_source="ilsvrc12_val_lmdb"):
    """reference: models/bvlc_alexnet/train_val.prototxt (conv→relu→norm→pool ordering)."""
    return _alex_like("AlexNet", batch, test_batch, classes, True, source, test_source)


def caffenet(batch=256, test_batch=50, classes=1000, source="ilsvrc12_train_lmdb",
             test_source="ilsvrc12_val_lmdb"):
    """reference: models/bvlc_reference_caffenet/train_val.prototxt (pool before norm)."""
    return _alex_like("CaffeNet", batch, test_batch, classes, False, source, test_source)


def alexnet_solver(net_path="train_val.prototxt", **over):
    """reference: models/bvlc_alexnet/solver.prototxt."""
    sp = P.SolverParameter(net=net_path, test_interval=1000, base_lr=0.005, lr_policy="step", gamma=0.1,
                           stepsize=15000, display=20, max_iter=45000, momentum=0.9, weight_decay=0.0005,
                           snapshot=10000, snapshot_prefix="caffe_alexnet_train", solver_mode="GPU")
    sp.test_iter = [1000]
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


def caffenet_solver(net_path="train_val.prototxt", **over):
    """reference: models/bvlc_reference_caffenet/solver.prototxt."""
    sp = alexnet_solver(net_path, base_lr=0.01, stepsize=100000, max_iter=450000, display=20,
                        snapshot_prefix="caffenet_train")
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


_INCEPTION = {  # name: (1x1, 3x3_reduce, 3x3, 5x5_reduce, 5x5, pool_proj)
    "3a": (64, 96, 128, 16, 32, 32), "3b": (128, 128, 192, 32, 96, 64),
    "4a": (192, 96, 208, 16, 48, 64), "4b": (160, 112, 224, 24, 64, 64),
    "4c": (128, 128, 256, 24, 64, 64), "4d": (112, 144, 288, 32, 64, 64),
    "4e": (256, 160, 320, 32, 128, 128), "5a": (256, 160, 320, 32, 128, 128),
    "5b": (384, 192, 384, 48, 128, 128),
}


def googlenet(batch=32, test_batch=50, classes=1000, source="ilsvrc12_train_lmdb",
              test_source="ilsvrc12_val_lmdb"):
    """reference: models/bvlc_googlenet/train_test.prototxt (59 conv, 5 IP, 9 concat, 3 losses
    weighted 0.3/0.3/1.0, dropout 0.7/0.7/0.4)."""
    b = NetBuilder("GoogleNet")
    b.data(source, batch, test_source, test_batch, crop=224, mirror=True, mean_values=(104, 117, 123))
    xf, bf = {"type": "xavier"}, _c(0.2)

    def cr(name, bottom, nout, k, stride=1, pad=0, relu_name=None):
        b.conv(name, bottom, nout, k, stride=stride, pad=pad, wf=xf, bf=bf)
        b.relu(relu_name or name.rsplit("/", 1)[0] + "/relu_" + name.rsplit("/", 1)[1], name)
        return name

    cr("conv1/7x7_s2", "data", 64, 7, 2, 3, "conv1/relu_7x7")
    b.pool("pool1/3x3_s2", "conv1/7x7_s2", "MAX", 3, 2)
    b.lrn("pool1/norm1", "pool1/3x3_s2")
    cr("conv2/3x3_reduce", "pool1/norm1", 64, 1)
    cr("conv2/3x3", "conv2/3x3_reduce", 192, 3, 1, 1)
    b.lrn("conv2/norm2", "conv2/3x3")
    x = b.pool("pool2/3x3_s2", "conv2/norm2", "MAX", 3, 2)

    def inception(tag, bottom):
        c1, r3, c3, r5, c5, pp = _INCEPTION[tag]
        p = f"inception_{tag}/"
        cr(p + "1x1", bottom, c1, 1)
        cr(p + "3x3_reduce", bottom, r3, 1)
        cr(p + "3x3", p + "3x3_reduce", c3, 3, 1, 1)
        cr(p + "5x5_reduce", bottom, r5, 1)
        cr(p + "5x5", p + "5x5_reduce", c5, 5, 1, 2)
        b.pool(p + "pool", bottom, "MAX", 3, 1, 1)
        cr(p + "pool_proj", p + "pool", pp, 1)
        return b.layer(p + "output", "CONCAT", (p + "1x1", p + "3x3", p + "5x5", p + "pool_proj"),
                       (p + "output",))

    def aux(idx, bottom):
        p = f"loss{idx}/"
        b.pool(p + "ave_pool", bottom, "AVE", 5, 3)
        cr(p + "conv", p + "ave_pool", 128, 1, relu_name=p + "relu_conv")
        b.fc(p + "fc", p + "conv", 1024, wf=xf, bf=bf)
        b.relu(p + "relu_fc", p + "fc")
        b.dropout(p + "drop_fc", p + "fc", 0.7)
        b.fc(p + "classifier", p + "fc", classes, wf=xf, bf=_c(0))
        # sic: the reference names both auxiliary tops ".../loss1"
        b.softmax_loss(p + "loss", p + "classifier", top=p + "loss1", weight=0.3)
        b.accuracy(p + "top-1", p + "classifier")
        b.accuracy(p + "top-5", p + "classifier", top_k=5)

    x = inception("3a", x)
    x = inception("3b", x)
    x = b.pool("pool3/3x3_s2", x, "MAX", 3, 2)
    x = inception("4a", x)
    aux(1, x)
    x = inception("4b", x)
    x = inception("4c", x)
    x = inception("4d", x)
    aux(2, x)
    x = inception("4e", x)
    x = b.pool("pool4/3x3_s2", x, "MAX", 3, 2)
    x = inception("5a", x)
    x = inception("5b", x)
    b.pool("pool5/7x7_s1", x, "AVE", 7, 1)
    b.dropout("pool5/drop_7x7_s1", "pool5/7x7_s1", 0.4)
    b.fc("loss3/classifier", "pool5/7x7_s1", classes, wf=xf, bf=_c(0))
    b.softmax_loss("loss3/loss3", "loss3/classifier", weight=1.0)
    b.accuracy("loss3/top-1", "loss3/classifier")
    b.accuracy("loss3/top-5", "loss3/classifier", top_k=5)
    return b.net


def googlenet_solver(net_path="train_test.prototxt", **over):
    """reference: models/bvlc_googlenet/quick_solver.prototxt."""
    sp = P.SolverParameter(net=net_path, test_interval=4000, test_initialization=True, display=40,
                           base_lr=0.01, lr_policy="poly", power=0.5, max_iter=2400000, momentum=0.9,
                           weight_decay=0.0002, snapshot=40000, snapshot_prefix="bvlc_googlenet_quick",
                           solver_mode="GPU")
    sp.test_iter = [1000]
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


def vgg16(batch=64, test_batch=50, classes=1000, source="ilsvrc12_train_lmdb",
          test_source="ilsvrc12_val_lmdb"):
    """VGG-16 (13 conv 3×3 + fc 25088→4096→4096→classes). Not in the reference tree; authored
    in the same V1 syntax as BASELINE.json config #5 asks (large fc6 stresses SFB vs dense)."""
    b = NetBuilder("VGG_ILSVRC_16_layers")
    b.data(source, batch, test_source, test_batch, crop=224, mirror=True, mean_values=(104, 117, 123))
    x = "data"
    cfg = [(64, 2), (128, 2), (256, 3), (512, 3), (512, 3)]
    for bi, (ch, n) in enumerate(cfg, 1):
        for j in range(1, n + 1):
            name = f"conv{bi}_{j}"
            b.conv(name, x, ch, 3, pad=1, wf={"type": "xavier"}, bf=_c(0))
            b.relu(f"relu{bi}_{j}", name)
            x = name
        x = b.pool(f"pool{bi}", x, "MAX", 2, 2)
    b.fc("fc6", x, 4096, wf=_g(0.005), bf=_c(0.1))
    b.relu("relu6", "fc6")
    b.dropout("drop6", "fc6", 0.5)
    b.fc("fc7", "fc6", 4096, wf=_g(0.005), bf=_c(0.1))
    b.relu("relu7", "fc7")
    b.dropout("drop7", "fc7", 0.5)
    b.fc("fc8", "fc7", classes, wf=_g(0.01), bf=_c(0))
    b.accuracy("accuracy", "fc8")
    b.softmax_loss("loss", "fc8")
    return b.net


def vgg16_solver(net_path="train_val.prototxt", **over):
    sp = P.SolverParameter(net=net_path, test_interval=1000, base_lr=0.01, lr_policy="step", gamma=0.1,
                           stepsize=100000, display=20, max_iter=370000, momentum=0.9, weight_decay=0.0005,
                           snapshot=10000, snapshot_prefix="vgg16_train", solver_mode="GPU")
    sp.test_iter = [1000]
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


def cifar10_quick_solver(net_path="cifar10_quick_train_test.prototxt", **over):
    """reference: examples/cifar10/cifar10_quick_solver.prototxt."""
    sp = P.SolverParameter(net=net_path, test_interval=500, base_lr=0.0007, momentum=0.9, weight_decay=0.004,
                           lr_policy="fixed", display=100, max_iter=4000, snapshot=4000,
                           snapshot_prefix="cifar10_quick", solver_mode="GPU")
    sp.test_iter = [100]
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


def cifar10_full_solver(net_path="cifar10_full_train_test.prototxt", **over):
    """reference: examples/cifar10/cifar10_full_solver.prototxt."""
    sp = P.SolverParameter(net=net_path, test_interval=1000, base_lr=0.001, momentum=0.9, weight_decay=0.004,
                           lr_policy="fixed", display=200, max_iter=60000, snapshot=10000,
                           snapshot_prefix="cifar10_full", solver_mode="GPU")
    sp.test_iter = [100]
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


# Follow-up solver stages the reference ships next to the main solver: the same run continued from the last snapshot
# with a lower learning rate (examples/cifar10/cifar10_{quick,full}_solver_lr{1,2}.prototxt), and GoogLeNet's long
# schedule (models/bvlc_googlenet/solver.prototxt).  name -> {file stem: overrides of the main solver}
SOLVER_STAGES = {
    "cifar10_quick": {"solver_lr1": dict(base_lr=0.0001, max_iter=5000, snapshot=5000)},
    "cifar10_full": {"solver_lr1": dict(base_lr=0.0001, max_iter=65000, snapshot=5000),
                     "solver_lr2": dict(base_lr=0.00001, max_iter=70000, snapshot=5000)},
    "googlenet": {"solver_long": dict(test_initialization=False, stepsize=320000, gamma=0.96, max_iter=10000000,
                                      snapshot_prefix="bvlc_googlenet")},
}


MODELS = {
    "lenet": (lenet, lenet_solver), "cifar10_quick": (cifar10_quick, cifar10_quick_solver),
    "cifar10_full": (cifar10_full, cifar10_full_solver),
    "alexnet": (alexnet, alexnet_solver), "caffenet": (caffenet, caffenet_solver),
    "googlenet": (googlenet, googlenet_solver), "vgg16": (vgg16, vgg16_solver),
}


def get_model(name: str, **kw):
    if name not in MODELS:
        raise KeyError(f"unknown model '{name}'; available: {sorted(MODELS)}")
    return MODELS[name][0](**kw)


def get_solver_param(name: str, net=None, **over):
    """SolverParameter for a zoo model with the net embedded (``net_param``)."""
    fn = MODELS[name][1]
    if fn is None:
        sp = P.SolverParameter(base_lr=0.001, momentum=0.9, weight_decay=0.004, lr_policy="fixed", display=100,
                               max_iter=4000, snapshot_prefix=name, solver_mode="GPU", test_interval=500)
        sp.test_iter = [100]
    else:
        sp = fn()
    sp.clear("net")
    sp.net_param = net if net is not None else get_model(name)
    for k, v in over.items():
        setattr(sp, k, v)
    return sp


_DATA_TYPES = ("DATA", "IMAGE_DATA", "WINDOW_DATA", "HDF5_DATA", "MEMORY_DATA", "DUMMY_DATA")
_LOSS_TYPES = ("SOFTMAX_LOSS", "EUCLIDEAN_LOSS", "HINGE_LOSS", "SIGMOID_CROSS_ENTROPY_LOSS", "MULTINOMIAL_LOGISTIC_LOSS",
               "INFOGAIN_LOSS", "CONTRASTIVE_LOSS")


def deploy(net, batch: int = 10, input_shape=None):
    """train_val -> deploy net, the transformation behind the reference's ``deploy.prototxt`` files
    (models/bvlc_alexnet/deploy.prototxt, examples/cifar10/cifar10_quick.prototxt): data layers become
    ``input: "data"`` + ``input_dim`` x 4, TEST/TRAIN-only layers and ACCURACY go away, the main SOFTMAX_LOSS becomes a
    SOFTMAX layer "prob", everything that does not feed "prob" (auxiliary heads, other losses, dropout keeps) is pruned.
    Layer names are kept, so a ``.caffemodel`` trained with the train_val net loads by name."""
    from ..net.net import filter_net
    layers = list(filter_net(net, P.NetState(phase=P.TEST)).layers)          # what a TEST-phase net would instantiate
    data_tops, shape, kept = [], input_shape, []
    for l in layers:
        t = l.enum_name("type")
        if t in _DATA_TYPES:
            if not data_tops:
                data_tops = list(l.top)
                if shape is None:
                    crop = int(l.transform_param.crop_size) if l.has("transform_param") else 0
                    src = (l.data_param.source if l.has("data_param") else "").lower()
                    if t == "MEMORY_DATA":
                        mp = l.memory_data_param
                        shape = (int(mp.channels), int(mp.height), int(mp.width))
                    elif "mnist" in src:
                        shape = (1, 28, 28)
                    elif "cifar" in src:
                        shape = (3, 32, 32)
                    else:
                        shape = (3, crop or 224, crop or 224)
            continue
        if t == "ACCURACY":
            continue
        kept.append(l)
    if not data_tops:
        raise ValueError("deploy(): the net has no data layer")
    # the main loss = the last loss layer with weight 1 (auxiliary heads carry 0.3 in GoogLeNet)
    main = None
    for l in kept:
        if l.enum_name("type") in _LOSS_TYPES and (not len(l.loss_weight) or float(l.loss_weight[0]) == 1.0):
            main = l
    out = P.NetParameter(name=net.name)
    out.input.append(data_tops[0])
    for d in (batch,) + tuple(shape):
        out.input_dim.append(int(d))
    body = [l for l in kept if l.enum_name("type") not in _LOSS_TYPES]
    if main is not None and main.enum_name("type") == "SOFTMAX_LOSS":
        prob = P.LayerParameter(name="prob", type="SOFTMAX")
        prob.bottom.append(main.bottom[0])
        prob.top.append("prob")
        body.append(prob)
        want = {"prob"}
    else:
        want = {main.bottom[0]} if main is not None else {body[-1].top[0]}
    needed = []
    for l in reversed(body):                                  # prune everything that does not reach the output
        if set(l.top) & want:
            needed.append(l)
            want |= set(l.bottom)
    for l in reversed(needed):
        out.layers.append(l.copy())
    return out


def write_zoo(out_dir: str, only=None):
    """Emit ``<model>/train_val.prototxt`` + ``solver.prototxt`` + ``deploy.prototxt`` for every (or the named) model."""
    import os
    for name, (net_fn, solver_fn) in MODELS.items():
        if only and name not in only:
            continue
        d = os.path.join(out_dir, name)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "train_val.prototxt"), "w") as f:
            f.write(to_text(net_fn()))
        with open(os.path.join(d, "deploy.prototxt"), "w") as f:
            f.write(to_text(deploy(net_fn())))
        if solver_fn is not None:
            sp = solver_fn(net_path=os.path.join(d, "train_val.prototxt"))
            with open(os.path.join(d, "solver.prototxt"), "w") as f:
                f.write(to_text(sp))
            for stem, over in SOLVER_STAGES.get(name, {}).items():
                st = solver_fn(net_path=os.path.join(d, "train_val.prototxt"), **over)
                if "stepsize" in over and "power" not in over:
                    st.clear("power")                       # the reference's long GoogLeNet solver: stepsize / gamma, no power
                with open(os.path.join(d, stem + ".prototxt"), "w") as f:
                    f.write(to_text(st))


def main(argv=None) -> int:
    """``python -m poseidon_b200.models.zoo --out models [--only alexnet,googlenet]``"""
    import argparse
    ap = argparse.ArgumentParser(description="write the model zoo as Caffe prototxt files")
    ap.add_argument("--out", default="models")
    ap.add_argument("--only", default="", help="comma separated subset of: " + ", ".join(MODELS))
    a = ap.parse_args(argv)
    only = [x for x in a.only.split(",") if x]
    for x in only:
        if x not in MODELS:
            ap.error(f"unknown model '{x}'")
    write_zoo(a.out, only or None)
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
