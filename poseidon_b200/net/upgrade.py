"""Upgrade deprecated net definitions to the current (V1 enum-typed) schema.

Two passes, both table-driven:
  1. V0 nets (``layers { layer { type: "conv" ... } }``) -> V1 LayerParameter, folding
     explicit "padding" layers into the consuming conv/pool layer.
  2. Legacy data-transform fields on DATA / IMAGE_DATA / WINDOW_DATA layers
     (scale / mean_file / crop_size / mirror) -> ``transform_param``.

reference: src/caffe/util/upgrade_proto.cpp:15-23 (NetNeedsUpgrade), :25-49
(UpgradeV0Net), :51-110 (padding fold), :112-452 (per-field moves), :454-506 (type map),
:508-580 (data-transform upgrade), :598-623 (UpgradeNetAsNeeded).
"""
from __future__ import annotations

import logging

from .. import proto as P

log = logging.getLogger("poseidon_b200")

V0_TYPE = {
    "accuracy": "ACCURACY", "bnll": "BNLL", "concat": "CONCAT", "conv": "CONVOLUTION",
    "data": "DATA", "dropout": "DROPOUT", "euclidean_loss": "EUCLIDEAN_LOSS",
    "flatten": "FLATTEN", "hdf5_data": "HDF5_DATA", "hdf5_output": "HDF5_OUTPUT",
    "im2col": "IM2COL", "images": "IMAGE_DATA", "infogain_loss": "INFOGAIN_LOSS",
    "innerproduct": "INNER_PRODUCT", "lrn": "LRN",
    "multinomial_logistic_loss": "MULTINOMIAL_LOGISTIC_LOSS", "pool": "POOLING",
    "relu": "RELU", "sigmoid": "SIGMOID", "softmax": "SOFTMAX",
    "softmax_loss": "SOFTMAX_LOSS", "split": "SPLIT", "tanh": "TANH",
    "window_data": "WINDOW_DATA",
}

# v0 field -> {v0 type: (v1 sub-message, v1 field)}
_MOVES = {
    "num_output": {"conv": ("convolution_param", "num_output"),
                   "innerproduct": ("inner_product_param", "num_output")},
    "biasterm": {"conv": ("convolution_param", "bias_term"),
                 "innerproduct": ("inner_product_param", "bias_term")},
    "weight_filler": {"conv": ("convolution_param", "weight_filler"),
                      "innerproduct": ("inner_product_param", "weight_filler")},
    "bias_filler": {"conv": ("convolution_param", "bias_filler"),
                    "innerproduct": ("inner_product_param", "bias_filler")},
    "pad": {"conv": ("convolution_param", "pad"), "pool": ("pooling_param", "pad")},
    "kernelsize": {"conv": ("convolution_param", "kernel_size"),
                   "pool": ("pooling_param", "kernel_size")},
    "group": {"conv": ("convolution_param", "group")},
    "stride": {"conv": ("convolution_param", "stride"), "pool": ("pooling_param", "stride")},
    "pool": {"pool": ("pooling_param", "pool")},
    "dropout_ratio": {"dropout": ("dropout_param", "dropout_ratio")},
    "local_size": {"lrn": ("lrn_param", "local_size")},
    "alpha": {"lrn": ("lrn_param", "alpha")},
    "beta": {"lrn": ("lrn_param", "beta")},
    "source": {"data": ("data_param", "source"), "hdf5_data": ("hdf5_data_param", "source"),
               "images": ("image_data_param", "source"),
               "window_data": ("window_data_param", "source"),
               "infogain_loss": ("infogain_loss_param", "source")},
    "scale": {"*": ("transform_param", "scale")},
    "meanfile": {"*": ("transform_param", "mean_file")},
    "cropsize": {"*": ("transform_param", "crop_size")},
    "mirror": {"*": ("transform_param", "mirror")},
    "batchsize": {"data": ("data_param", "batch_size"),
                  "hdf5_data": ("hdf5_data_param", "batch_size"),
                  "images": ("image_data_param", "batch_size"),
                  "window_data": ("window_data_param", "batch_size")},
    "rand_skip": {"data": ("data_param", "rand_skip"),
                  "images": ("image_data_param", "rand_skip")},
    "shuffle_images": {"images": ("image_data_param", "shuffle")},
    "new_height": {"images": ("image_data_param", "new_height")},
    "new_width": {"images": ("image_data_param", "new_width")},
    "concat_dim": {"concat": ("concat_param", "concat_dim")},
    "det_fg_threshold": {"window_data": ("window_data_param", "fg_threshold")},
    "det_bg_threshold": {"window_data": ("window_data_param", "bg_threshold")},
    "det_fg_fraction": {"window_data": ("window_data_param", "fg_fraction")},
    "det_context_pad": {"window_data": ("window_data_param", "context_pad")},
    "det_crop_mode": {"window_data": ("window_data_param", "crop_mode")},
    "hdf5_output_param": {"hdf5_output": ("hdf5_output_param", None)},
}


def net_needs_v0_upgrade(net) -> bool:
    return any(l.has("layer") for l in net.layers)


def _fold_padding_layers(net):
    out = net.copy()
    out.clear("layers")
    last_top = {name: -1 for name in net.input}
    for i, conn in enumerate(net.layers):
        v0 = conn.layer
        if v0.type != "padding":
            out.layers.append(conn.copy())
        for j, bname in enumerate(conn.bottom):
            if bname not in last_top:
                raise ValueError(f"unknown blob input {bname} to layer {i}")
            src_idx = last_top[bname]
            if src_idx < 0:
                continue
            src = net.layers[src_idx]
            if src.layer.type == "padding":
                if v0.type not in ("conv", "pool"):
                    raise ValueError("padding layer feeds non conv/pool layer " + str(v0.type))
                if len(conn.bottom) != 1 or len(src.bottom) != 1 or len(src.top) != 1:
                    raise ValueError("padding layers must be single-input single-output")
                tgt = out.layers[-1]
                tgt.mutable("layer").pad = src.layer.pad
                tgt.bottom[j] = src.bottom[0]
        for tname in conn.top:
            last_top[tname] = i
    return out


def _upgrade_layer(conn):
    new = P.LayerParameter()
    new.bottom = list(conn.bottom)
    new.top = list(conn.top)
    ok = True
    if not conn.has("layer"):
        return new, ok
    v0 = conn.layer
    typ = v0.type
    if v0.has("name"):
        new.name = v0.name
    if v0.has("type"):
        if typ in V0_TYPE:
            new.type = V0_TYPE[typ]
        else:
            log.error("unknown V0 layer type %s", typ)
            new.type = "NONE"
            ok = False
    for b in v0.blobs:
        new.blobs.append(b.copy())
    new.blobs_lr = v0.blobs_lr
    new.weight_decay = v0.weight_decay
    for fname, table in _MOVES.items():
        if not v0.has(fname):
            continue
        dest = table.get(typ) or table.get("*")
        if dest is None:
            log.error("unknown parameter %s for V0 layer type %s", fname, typ)
            ok = False
            continue
        sub, field = dest
        if field is None:
            new.mutable(sub).CopyFrom(getattr(v0, fname))
        else:
            val = getattr(v0, fname)
            setattr(new.mutable(sub), field, val.copy() if isinstance(val, P.Message) else val)
    for fname in ("new_num", "new_channels"):
        if v0.has(fname):
            log.error("unknown parameter %s for V0 layer type %s", fname, typ)
            ok = False
    return new, ok


def upgrade_v0_net(net):
    folded = _fold_padding_layers(net)
    out = P.NetParameter()
    if folded.has("name"):
        out.name = folded.name
    ok = True
    for conn in folded.layers:
        new, good = _upgrade_layer(conn)
        ok &= good
        out.layers.append(new)
    out.input = list(folded.input)
    out.input_dim = list(folded.input_dim)
    if folded.has("force_backward"):
        out.force_backward = folded.force_backward
    return out, ok


_XFORM_LAYERS = {
    P.LayerType["DATA"]: "data_param",
    P.LayerType["IMAGE_DATA"]: "image_data_param",
    P.LayerType["WINDOW_DATA"]: "window_data_param",
}
_XFORM_FIELDS = ("scale", "mean_file", "crop_size", "mirror")


def net_needs_data_upgrade(net) -> bool:
    for l in net.layers:
        sub = _XFORM_LAYERS.get(l.type)
        if sub and l.has(sub) and any(getattr(l, sub).has(f) for f in _XFORM_FIELDS):
            return True
    return False


def upgrade_data_transformation(net) -> None:
    for l in net.layers:
        sub = _XFORM_LAYERS.get(l.type)
        if not sub or not l.has(sub):
            continue
        lp = getattr(l, sub)
        for f in _XFORM_FIELDS:
            if lp.has(f):
                setattr(l.mutable("transform_param"), f, getattr(lp, f))
                lp.clear(f)


def upgrade_net_as_needed(net, source: str = "<net>"):
    if net_needs_v0_upgrade(net):
        log.warning("upgrading deprecated V0LayerParameter net: %s", source)
        net, ok = upgrade_v0_net(net)
        if not ok:
            log.error("problems upgrading V0 net %s; continuing", source)
    if net_needs_data_upgrade(net):
        log.warning("upgrading deprecated data transformation params: %s", source)
        upgrade_data_transformation(net)
    return net
