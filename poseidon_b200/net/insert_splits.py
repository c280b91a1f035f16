"""Explicit SPLIT-layer insertion for blobs consumed by more than one layer.

The executor itself never needs Split layers (autograd accumulates fan-out gradients), but the transformation is
part of the reference's observable behaviour — ``Net::ToProto`` output contains the inserted layers and their
naming scheme — so it is provided as a pure NetParameter -> NetParameter function.

reference: src/caffe/util/insert_splits.cpp:12-144 (InsertSplits / ConfigureSplitLayer / SplitLayerName /
SplitBlobName: ``<blob>_<layer>_<top_idx>_split`` and ``..._split_<k>``).
"""
from __future__ import annotations

from typing import Dict, Tuple

from .. import proto as P


def split_layer_name(layer_name: str, blob_name: str, blob_idx: int) -> str:
    return f"{blob_name}_{layer_name}_{blob_idx}_split"


def split_blob_name(layer_name: str, blob_name: str, blob_idx: int, split_idx: int) -> str:
    return f"{blob_name}_{layer_name}_{blob_idx}_split_{split_idx}"


def insert_splits(param):
    out = param.copy()
    out.clear("layers")
    blob_to_top: Dict[str, Tuple[int, int]] = {}          # blob -> (layer idx, top idx) of its latest producer
    bottom_to_top: Dict[Tuple[int, int], Tuple[int, int]] = {}
    top_count: Dict[Tuple[int, int], int] = {}
    top_loss_weight: Dict[Tuple[int, int], float] = {}
    layer_names = {-1: "input"}
    for i, name in enumerate(param.input):
        blob_to_top[name] = (-1, i)
    for i, lp in enumerate(param.layers):
        layer_names[i] = lp.name
        for j, b in enumerate(lp.bottom):
            if b not in blob_to_top:
                raise ValueError(f"Unknown blob input {b} to layer {lp.name}")
            src = blob_to_top[b]
            bottom_to_top[(i, j)] = src
            top_count[src] = top_count.get(src, 0) + 1
        for j, t in enumerate(lp.top):
            blob_to_top[t] = (i, j)
        for j, w in enumerate(lp.loss_weight):
            if j < len(lp.top) and w:
                key = (i, j)
                top_loss_weight[key] = float(w)
                top_count[key] = top_count.get(key, 0) + 1      # the loss itself is a consumer

    def make_split(layer_idx, top_idx, blob, loss_weight=0.0):
        n = top_count[(layer_idx, top_idx)]
        sl = P.LayerParameter(name=split_layer_name(layer_names[layer_idx], blob, top_idx), type="SPLIT")
        sl.bottom = [blob]
        sl.top = [split_blob_name(layer_names[layer_idx], blob, top_idx, k) for k in range(n)]
        if loss_weight:
            sl.loss_weight = [loss_weight] + [0.0] * (n - 1)
        return sl

    for i, name in enumerate(param.input):
        if top_count.get((-1, i), 0) > 1:
            out.layers.append(make_split(-1, i, name))
    top_used: Dict[Tuple[int, int], int] = {}
    for i, lp in enumerate(param.layers):
        nl = lp.copy()
        for j, b in enumerate(lp.bottom):
            src = bottom_to_top[(i, j)]
            if top_count.get(src, 0) > 1:
                k = top_used.get(src, 0)
                if src in top_loss_weight and k == 0:
                    k = 1                     # branch 0 carries the loss weight
                    top_used[src] = 1
                nl.bottom[j] = split_blob_name(layer_names[src[0]], b, src[1], k)
                top_used[src] = top_used.get(src, 0) + 1 if src not in top_loss_weight or k > 1 else k + 1
        out.layers.append(nl)
        for j, t in enumerate(lp.top):
            if top_count.get((i, j), 0) > 1:
                lw = top_loss_weight.get((i, j), 0.0)
                if lw:
                    nl.loss_weight = [0.0 if q == j else float(w) for q, w in enumerate(list(nl.loss_weight))]
                out.layers.append(make_split(i, j, t, lw))
    return out
