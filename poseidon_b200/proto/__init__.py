"""Caffe-compatible proto messages + file IO (text ``.prototxt`` / binary ``.caffemodel``).

reference: src/caffe/util/io.cpp:31-80 (ReadProtoFromTextFile, ReadProtoFromBinaryFile,
WriteProtoToTextFile, WriteProtoToBinaryFile).
"""
from __future__ import annotations

import os

import numpy as np

from . import schema
from .message import DecodeError, Message, ParseError, get_class, parse_text, to_text

SVProto = get_class("SVProto")
BlobProto = get_class("BlobProto")
BlobProtoVector = get_class("BlobProtoVector")
Datum = get_class("Datum")
FillerParameter = get_class("FillerParameter")
NetParameter = get_class("NetParameter")
SolverParameter = get_class("SolverParameter")
SolverState = get_class("SolverState")
NetState = get_class("NetState")
NetStateRule = get_class("NetStateRule")
LayerParameter = get_class("LayerParameter")
LayerPSTablePair = get_class("LayerPSTablePair")
V0LayerParameter = get_class("V0LayerParameter")
TransformationParameter = get_class("TransformationParameter")

TRAIN, TEST = schema.ENUMS["Phase"]["TRAIN"], schema.ENUMS["Phase"]["TEST"]
LayerType = schema.ENUMS["LayerType"]
LayerTypeName = {v: k for k, v in LayerType.items()}


def read_text(path: str, cls):
    with open(path, "r") as f:
        return parse_text(f.read(), cls)


def write_text(path: str, msg: Message) -> None:
    with open(path, "w") as f:
        f.write(to_text(msg))


def read_binary(path: str, cls):
    with open(path, "rb") as f:
        data = f.read()
    return cls.FromString(data)


def write_binary(path: str, msg: Message) -> None:
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(msg.SerializeToString())
    os.replace(tmp, path)


def read_net(path: str):
    """Read a NetParameter from prototxt (text) or caffemodel (binary) and upgrade it."""
    from ..net.upgrade import upgrade_net_as_needed

    with open(path, "rb") as f:
        data = f.read()
    try:
        net = parse_text(data.decode("utf-8"), NetParameter)
    except (UnicodeDecodeError, ValueError):
        net = NetParameter.FromString(data)
    return upgrade_net_as_needed(net)


def read_solver(path: str):
    return read_text(path, SolverParameter)


def blob_to_array(blob) -> np.ndarray:
    shape = (blob.num, blob.channels, blob.height, blob.width)
    return np.asarray(blob.data, dtype=np.float32).reshape(shape)


def array_to_blob(arr, diff=None, blob_mode=None, global_id=None):
    """4-D NCHW blob; lower-rank arrays are right-aligned Caffe style (1,1,N,K)/(1,1,1,N)."""
    arr = np.asarray(arr, dtype=np.float32)
    shape = (1,) * (4 - arr.ndim) + tuple(arr.shape)
    b = BlobProto(num=shape[0], channels=shape[1], height=shape[2], width=shape[3])
    b.data = arr.reshape(-1)
    if diff is not None:
        b.diff = np.asarray(diff, dtype=np.float32).reshape(-1)
    if blob_mode is not None:
        b.blob_mode = blob_mode
    if global_id is not None:
        b.global_id = global_id
    return b
