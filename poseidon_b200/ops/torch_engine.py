"""The "torch" engine: CPU and vendor-library (cuDNN/cuBLAS via PyTorch) execution of
every op.  This is the *baseline* engine — the product is :mod:`.sm100`.
"""
from __future__ import annotations

import torch

from . import reference as R

max_pool = R.max_pool
ave_pool = R.ave_pool
stochastic_pool = R.stochastic_pool
lrn_across = R.lrn_across
lrn_within = R.lrn_within
relu = R.relu
softmax = R.softmax
softmax_loss = R.softmax_loss
sigmoid = torch.sigmoid
tanh = torch.tanh
absval = torch.abs
bnll = R.bnll
power = R.power
eltwise = R.eltwise
mvn = R.mvn


def threshold(x, t):
    return (x > t).to(x.dtype)


def conv2d(x, w, b, stride, pad, groups, relu_slope=None, layer=None):
    y = R.conv2d(x, w, b, stride, pad, groups)
    return y if relu_slope is None else R.relu(y, relu_slope)


def inner_product(x, w, b, relu=False, layer=None):
    sfb = getattr(layer, "sfb", None) if layer is not None else None
    if sfb is not None and torch.is_grad_enabled():
        return sfb.apply(layer, x, w, b, relu)
    y = R.inner_product(x, w, b)
    return torch.relu(y) if relu else y


def dropout(x, ratio, train):
    return R.dropout(x, ratio, train)


def concat(xs, dim, layer=None):
    return torch.cat(list(xs), dim=dim)


def transform(transformer, x, out_dtype, first_conv=None):
    y = transformer(x, out_dtype)
    if y.is_cuda and out_dtype != torch.float32:
        y = y.contiguous(memory_format=torch.channels_last)     # cuDNN's tensor-core friendly layout
    return y
