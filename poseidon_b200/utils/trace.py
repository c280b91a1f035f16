"""NVTX ranges for layers and communication launches (``POSEIDON_NVTX=1``).

The reference has no tracing hooks at all; here every layer forward, every bucket launch and the
optimizer step can be bracketed with NVTX ranges so that nsys / ncu timelines name them.  Off by default: the
range push/pop costs a few hundred nanoseconds per call on a launch-bound net.
"""
from __future__ import annotations

import contextlib
import os

import torch

ENABLED = os.environ.get("POSEIDON_NVTX", "0") not in ("0", "", "false")


@contextlib.contextmanager
def nvtx_range(name: str):
    if ENABLED and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


def annotate_net(net) -> int:
    """Wrap every layer's forward in an NVTX range named ``fwd/<layer>`` (idempotent).  Returns #layers wrapped."""
    if not ENABLED:
        return 0
    n = 0
    for layer in net.layers:
        if getattr(layer, "_nvtx_wrapped", False):
            continue
        fwd = layer.forward

        def wrapped(*a, _f=fwd, _n=f"fwd/{layer.layer_name}", **kw):
            with nvtx_range(_n):
                return _f(*a, **kw)
        layer.forward = wrapped
        layer._nvtx_wrapped = True
        n += 1
    return n
