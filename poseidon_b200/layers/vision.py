"""Vision layers: Convolution, Pooling, LRN, Im2col.

reference: include/caffe/vision_layers.hpp:36 (Convolution), :132 (Im2col), :171 (LRN),
:215 (Pooling); the cuDNN variants (:294, :343) collapse into the engine switch.
"""
from __future__ import annotations


from .. import ops
from .base import Layer, hw_param, register


@register("CONVOLUTION")
class ConvolutionLayer(Layer):
    """Grouped 2-D convolution, weight (Cout, Cin/g, kh, kw), bias (Cout,).
    reference: src/caffe/layers/conv_layer.cpp:12-155 (setup/reshape), conv_layer.cu:13-119."""
    min_bottoms = 1
    min_tops = 1
    fused_relu_slope = None   # set by the net's fusion pass (sm100 engine epilogue)

    def setup(self, bottom_shapes):
        cp = self.lp.convolution_param
        self.kernel = hw_param(cp, "kernel_size", "kernel_h", "kernel_w")
        self.pad = hw_param(cp, "pad", "pad_h", "pad_w", 0)
        self.stride = hw_param(cp, "stride", "stride_h", "stride_w", 1)
        self.group = int(cp.group)
        self.num_output = int(cp.num_output)
        self.bias_term = bool(cp.bias_term)
        n, c, h, w = bottom_shapes[0]
        self.in_hw = (h, w)
        for s in bottom_shapes[1:]:
            if tuple(s) != tuple(bottom_shapes[0]):
                raise ValueError("all conv bottoms must have the same shape")
        if c % self.group or self.num_output % self.group:
            raise ValueError("channels and num_output must be multiples of group")
        self.add_blob("weight", (self.num_output, c // self.group) + self.kernel, cp.weight_filler
                      if cp.has("weight_filler") else None)
        if self.bias_term:
            self.add_blob("bias", (self.num_output,), cp.bias_filler if cp.has("bias_filler") else None)
        oh = (h + 2 * self.pad[0] - self.kernel[0]) // self.stride[0] + 1
        ow = (w + 2 * self.pad[1] - self.kernel[1]) // self.stride[1] + 1
        return [(n, self.num_output, oh, ow)] * len(bottom_shapes)

    def forward(self, *bottoms):
        k = ops.get(self.ctx)
        b = self.bias if self.bias_term else None
        return tuple(k.conv2d(x, self.weight, b, self.stride, self.pad, self.group,
                              relu_slope=self.fused_relu_slope, layer=self) for x in bottoms)


@register("POOLING")
class PoolingLayer(Layer):
    """MAX (optional mask top) / AVE / STOCHASTIC pooling with Caffe's ceil output size.
    reference: src/caffe/layers/pooling_layer.cpp:18-103, pooling_layer.cu:12-379."""
    exact_bottoms = 1
    min_tops = 1
    max_tops = 2

    def setup(self, bottom_shapes):
        pp = self.lp.pooling_param
        self.kernel = hw_param(pp, "kernel_size", "kernel_h", "kernel_w")
        self.pad = hw_param(pp, "pad", "pad_h", "pad_w", 0)
        self.stride = hw_param(pp, "stride", "stride_h", "stride_w", 1)
        self.method = pp.enum_name("pool")
        if self.pad != (0, 0):
            if self.method == "STOCHASTIC":
                raise ValueError("padding implemented only for average and max pooling")
            if self.pad[0] >= self.kernel[0] or self.pad[1] >= self.kernel[1]:
                raise ValueError("pad must be smaller than kernel")
        n, c, h, w = bottom_shapes[0]
        oh = ops.reference.pool_out_size(h, self.kernel[0], self.stride[0], self.pad[0])
        ow = ops.reference.pool_out_size(w, self.kernel[1], self.stride[1], self.pad[1])
        return [(n, c, oh, ow)] * 2

    def forward(self, x):
        k = ops.get(self.ctx)
        if self.method == "MAX":
            if self.n_tops == 2:
                y, mask = k.max_pool(x, self.kernel, self.stride, self.pad, return_mask=True)
                return y, mask
            return (k.max_pool(x, self.kernel, self.stride, self.pad, **getattr(self, "engine_kw", {})),)
        if self.method == "AVE":
            return (k.ave_pool(x, self.kernel, self.stride, self.pad),)
        return (k.stochastic_pool(x, self.kernel, self.stride, self.ctx.train),)


@register("LRN")
class LRNLayer(Layer):
    """Local response normalisation, across or within channels.
    reference: src/caffe/layers/lrn_layer.cpp:20-165, lrn_layer.cu:10-190."""
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        p = self.lp.lrn_param
        self.size = int(p.local_size)
        if self.size % 2 == 0:
            raise ValueError("LRN only supports odd values for local_size")
        self.alpha, self.beta = float(p.alpha), float(p.beta)
        self.region = p.enum_name("norm_region")
        return [tuple(bottom_shapes[0])]

    def forward(self, x):
        k = ops.get(self.ctx)
        if self.region == "ACROSS_CHANNELS":
            return (k.lrn_across(x, self.size, self.alpha, self.beta, **getattr(self, "engine_kw", {})),)
        return (k.lrn_within(x, self.size, self.alpha, self.beta),)


@register("IM2COL")
class Im2colLayer(Layer):
    """reference: src/caffe/layers/im2col_layer.cpp, util/im2col.cu:12-40."""
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        cp = self.lp.convolution_param
        self.kernel = hw_param(cp, "kernel_size", "kernel_h", "kernel_w")
        self.pad = hw_param(cp, "pad", "pad_h", "pad_w", 0)
        self.stride = hw_param(cp, "stride", "stride_h", "stride_w", 1)
        n, c, h, w = bottom_shapes[0]
        oh = (h + 2 * self.pad[0] - self.kernel[0]) // self.stride[0] + 1
        ow = (w + 2 * self.pad[1] - self.kernel[1]) // self.stride[1] + 1
        return [(n, c * self.kernel[0] * self.kernel[1], oh, ow)]

    def forward(self, x):
        return (ops.reference.im2col(x, self.kernel, self.stride, self.pad),)
