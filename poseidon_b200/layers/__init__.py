"""Layer registry: importing this package registers every Caffe layer type."""
from .base import LAYER_REGISTRY, Layer, NetContext, fill, set_filler_seed  # noqa: F401
from . import common, data, loss, neuron, vision  # noqa: F401


def create_layer(lp, ctx):
    """reference: src/caffe/layer_factory.cpp:177-261 (GetLayer switch on LayerParameter.type)."""
    from .. import proto as P
    tname = P.LayerTypeName.get(lp.type)
    if tname is None or tname == "NONE":
        raise ValueError(f"Layer '{lp.name}' has unspecified or unknown type {lp.type}")
    cls = LAYER_REGISTRY.get(tname)
    if cls is None:
        raise ValueError(f"Unknown layer type: {tname}")
    return cls(lp, ctx)
