"""Common layers: InnerProduct, Softmax, Concat, Slice, Split, Flatten, Eltwise, ArgMax,
MVN, Silence.

reference: include/caffe/common_layers.hpp:29 (ArgMax), :82 (Concat), :164 (Eltwise),
:209 (Flatten), :260 (InnerProduct), :303 (MVN), :337 (Silence), :369 (Softmax),
:436 (Split), :469 (Slice).
"""
from __future__ import annotations

import torch

from .. import ops
from .base import Layer, register


@register("INNER_PRODUCT")
class InnerProductLayer(Layer):
    """top = X·Wᵀ + b; weight (N, K) stored as blob (1,1,N,K), bias (N,) as (1,1,1,N).

    Under sufficient-factor broadcasting the dense weight gradient is never formed: the
    layer hands (u=top_diff, v=bottom) to the SFB engine instead (``sfb`` attribute set by
    the parallel engine).
    reference: src/caffe/layers/inner_product_layer.cpp:14-135, inner_product_layer.cu:13-64."""
    exact_bottoms = 1
    exact_tops = 1
    fused_relu = False
    sfb = None

    def setup(self, bottom_shapes):
        p = self.lp.inner_product_param
        self.num_output = int(p.num_output)
        self.bias_term = bool(p.bias_term)
        n = bottom_shapes[0][0]
        k = 1
        for d in bottom_shapes[0][1:]:
            k *= d
        self.K = k
        self.add_blob("weight", (self.num_output, k), p.weight_filler if p.has("weight_filler") else None)
        if self.bias_term:
            self.add_blob("bias", (self.num_output,), p.bias_filler if p.has("bias_filler") else None)
        return [(n, self.num_output, 1, 1)]

    def forward(self, x):
        k = ops.get(self.ctx)
        b = self.bias if self.bias_term else None
        y = k.inner_product(x, self.weight, b, relu=self.fused_relu, layer=self)
        return (y.reshape(y.shape[0], self.num_output, 1, 1),)


@register("SOFTMAX")
class SoftmaxLayer(Layer):
    """reference: src/caffe/layers/softmax_layer.cu:88-149."""
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        return [tuple(bottom_shapes[0])]

    def forward(self, x):
        return (ops.get(self.ctx).softmax(x),)


@register("CONCAT")
class ConcatLayer(Layer):
    """reference: src/caffe/layers/concat_layer.cpp, concat_layer.cu:10-72 (dim 0 or 1)."""
    min_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        self.dim = int(self.lp.concat_param.concat_dim)
        if self.dim not in (0, 1):
            raise ValueError("concat_dim must be 0 or 1")
        out = list(bottom_shapes[0])
        for s in bottom_shapes[1:]:
            for d in range(4):
                if d != self.dim and s[d] != out[d]:
                    raise ValueError("concat bottoms disagree on non-concat dims")
            out[self.dim] += s[self.dim]
        return [tuple(out)]

    _slab = None          # sm100 engine, zero-copy concat: the output buffer the producers write their slices into
    _slab_channels = 0

    def slab_view(self, offset: int, n: int, h: int, w: int, device):
        """Channel slice [offset, offset + c) of this pass's output slab (allocated by the first producer that asks);
        planned by net/fusion.py for bottoms produced by convolution kernels."""
        import torch
        if self._slab is None or tuple(self._slab.shape) != (n, self._slab_channels, h, w) or self._slab.device != device:
            # (memory_format in the factory: empty(...).contiguous(channels_last) would launch a copy of the whole slab)
            self._slab = torch.empty((n, self._slab_channels, h, w), device=device, dtype=torch.bfloat16,
                                     memory_format=torch.channels_last)
        c = self._slab_slices[offset]
        slab = self._slab
        # an ALIAS of the slice, not an autograd view of the slab: the producers' kernels write disjoint slices of one
        # buffer, which view tracking would flag as in-place modification of a shared base
        return torch.empty(0, dtype=slab.dtype, device=slab.device).set_(
            slab.untyped_storage(), slab.storage_offset() + offset, (n, c, h, w), slab.stride())

    def forward(self, *xs):
        return (ops.get(self.ctx).concat(xs, self.dim, layer=self),)


@register("SLICE")
class SliceLayer(Layer):
    """reference: src/caffe/layers/slice_layer.cpp:10-60."""
    exact_bottoms = 1
    min_tops = 2

    def setup(self, bottom_shapes):
        p = self.lp.slice_param
        self.dim = int(p.slice_dim)
        if self.dim not in (0, 1):
            raise ValueError("slice_dim must be 0 or 1")
        total = bottom_shapes[0][self.dim]
        pts = [int(x) for x in p.slice_point]
        if pts:
            if len(pts) != self.n_tops - 1:
                raise ValueError("need n_tops-1 slice points")
            edges = [0] + pts + [total]
            self.sizes = [b - a for a, b in zip(edges[:-1], edges[1:])]
            if any(s <= 0 for s in self.sizes):
                raise ValueError("slice points must be strictly increasing")
        else:
            if total % self.n_tops:
                raise ValueError("slice dim not divisible by number of tops")
            self.sizes = [total // self.n_tops] * self.n_tops
        shapes = []
        for s in self.sizes:
            sh = list(bottom_shapes[0])
            sh[self.dim] = s
            shapes.append(tuple(sh))
        return shapes

    def forward(self, x):
        return tuple(torch.split(x, self.sizes, dim=self.dim))


@register("SPLIT")
class SplitLayer(Layer):
    """Fan-out is implicit under autograd; SPLIT layers in existing nets are identities.
    reference: src/caffe/layers/split_layer.cpp, util/insert_splits.cpp:12-144."""
    exact_bottoms = 1
    min_tops = 1

    def setup(self, bottom_shapes):
        return [tuple(bottom_shapes[0])] * self.n_tops

    def forward(self, x):
        return (x,) * self.n_tops


@register("FLATTEN")
class FlattenLayer(Layer):
    """reference: src/caffe/layers/flatten_layer.cpp."""
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        n, c, h, w = bottom_shapes[0]
        return [(n, c * h * w, 1, 1)]

    def forward(self, x):
        if x.dim() == 4 and not x.is_contiguous():
            x = x.contiguous()   # Caffe's flatten order is C,H,W
        return (x.reshape(x.shape[0], -1, 1, 1),)


@register("ELTWISE")
class EltwiseLayer(Layer):
    """reference: src/caffe/layers/eltwise_layer.cpp:12-80."""
    min_bottoms = 2
    exact_tops = 1

    def setup(self, bottom_shapes):
        p = self.lp.eltwise_param
        self.op = p.enum_name("operation")
        self.coeffs = [float(c) for c in p.coeff]
        if self.coeffs and self.op != "SUM":
            raise ValueError("Eltwise layer only takes coefficients for summation")
        if self.coeffs and len(self.coeffs) != len(bottom_shapes):
            raise ValueError("Eltwise coeff count must match bottoms")
        for s in bottom_shapes[1:]:
            if tuple(s) != tuple(bottom_shapes[0]):
                raise ValueError("Eltwise bottoms must have equal shapes")
        return [tuple(bottom_shapes[0])]

    def forward(self, *xs):
        return (ops.get(self.ctx).eltwise(xs, self.op, self.coeffs or None),)


@register("ARGMAX")
class ArgMaxLayer(Layer):
    """reference: src/caffe/layers/argmax_layer.cpp:10-60."""
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        p = self.lp.argmax_param
        self.top_k, self.out_max_val = int(p.top_k), bool(p.out_max_val)
        return [(bottom_shapes[0][0], 2 if self.out_max_val else 1, self.top_k, 1)]

    def forward(self, x):
        return (ops.reference.argmax(x, self.top_k, self.out_max_val),)


@register("MVN")
class MVNLayer(Layer):
    """reference: src/caffe/layers/mvn_layer.cpp:39-69."""
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        p = self.lp.mvn_param
        self.nv, self.ac = bool(p.normalize_variance), bool(p.across_channels)
        return [tuple(bottom_shapes[0])]

    def forward(self, x):
        return (ops.get(self.ctx).mvn(x, self.nv, self.ac),)


@register("SILENCE")
class SilenceLayer(Layer):
    """Consumes blobs so they don't become net outputs. reference: silence_layer.cpp."""
    min_bottoms = 1
    exact_tops = 0

    def setup(self, bottom_shapes):
        return []

    def forward(self, *xs):
        return ()
