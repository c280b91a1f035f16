"""Neuron (elementwise) layers.

reference: include/caffe/neuron_layers.hpp:25 (Neuron), :50 (AbsVal), :113 (BNLL),
:163 (Dropout), :225 (Power), :301 (ReLU), :406 (Sigmoid), :493 (TanH), :578 (Threshold);
cuDNN variants (:372, :459, :548) collapse into the engine switch.
"""
from __future__ import annotations


from .. import ops
from .base import Layer, register


class NeuronLayer(Layer):
    exact_bottoms = 1
    exact_tops = 1

    def setup(self, bottom_shapes):
        self.configure()
        return [tuple(bottom_shapes[0])]

    def configure(self):
        pass


@register("RELU")
class ReLULayer(NeuronLayer):
    """reference: src/caffe/layers/relu_layer.cu:10-59 (leaky via negative_slope)."""

    def configure(self):
        self.slope = float(self.lp.relu_param.negative_slope)

    def forward(self, x):
        return (ops.get(self.ctx).relu(x, self.slope),)


@register("SIGMOID")
class SigmoidLayer(NeuronLayer):
    """reference: src/caffe/layers/sigmoid_layer.cu."""

    def forward(self, x):
        return (ops.get(self.ctx).sigmoid(x),)


@register("TANH")
class TanHLayer(NeuronLayer):
    """reference: src/caffe/layers/tanh_layer.cu."""

    def forward(self, x):
        return (ops.get(self.ctx).tanh(x),)


@register("ABSVAL")
class AbsValLayer(NeuronLayer):
    """reference: src/caffe/layers/absval_layer.cpp."""

    def forward(self, x):
        return (ops.get(self.ctx).absval(x),)


@register("BNLL")
class BNLLLayer(NeuronLayer):
    """reference: src/caffe/layers/bnll_layer.cpp:19-20."""

    def forward(self, x):
        return (ops.get(self.ctx).bnll(x),)


@register("POWER")
class PowerLayer(NeuronLayer):
    """reference: src/caffe/layers/power_layer.cpp:22-45."""

    def configure(self):
        p = self.lp.power_param
        self.power, self.scale, self.shift = float(p.power), float(p.scale), float(p.shift)

    def forward(self, x):
        return (ops.get(self.ctx).power(x, self.power, self.scale, self.shift),)


@register("THRESHOLD")
class ThresholdLayer(NeuronLayer):
    """x > t ? 1 : 0 (no gradient). Unreachable from the reference factory but
    implemented there; we make it reachable. reference: src/caffe/layers/threshold_layer.cpp."""

    def configure(self):
        self.threshold = float(self.lp.threshold_param.threshold)

    def forward(self, x):
        return (ops.get(self.ctx).threshold(x, self.threshold),)


@register("DROPOUT")
class DropoutLayer(NeuronLayer):
    """reference: src/caffe/layers/dropout_layer.cpp:17-47, dropout_layer.cu:15-70."""

    def configure(self):
        self.ratio = float(self.lp.dropout_param.dropout_ratio)
        if not (0.0 < self.ratio < 1.0):
            raise ValueError("dropout_ratio must be in (0, 1)")

    def forward(self, x):
        return (ops.get(self.ctx).dropout(x, self.ratio, self.ctx.train),)
