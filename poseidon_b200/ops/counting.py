"""Launch accounting for the sm100 engine: every call into ``torch.ops.poseidon`` is tallied (weighted by the
number of CUDA kernels the op launches) so benchmarks can report how many of OUR kernels ran in a region."""
from __future__ import annotations

from collections import Counter

_counts: Counter = Counter()
# kernels launched per op call where it is not 1
_WEIGHT = {"colsum": 2, "conv_fprop": 1, "conv_dgrad": 1, "conv_wgrad": 1, "allreduce_sgd": 1}


def record(name: str, n: int = 1):
    _counts[name] += n * _WEIGHT.get(name, 1)


def reset():
    _counts.clear()


def total() -> int:
    return int(sum(_counts.values()))


def by_op():
    return dict(_counts)


class CountingOps:
    """Proxy around ``torch.ops.poseidon`` that records every op call."""

    def __init__(self, ops):
        self._ops = ops
        self._cache = {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            op = getattr(self._ops, name)

            def fn(*a, _op=op, _name=name, **kw):
                _counts[_name] += _WEIGHT.get(_name, 1)
                return _op(*a, **kw)
            self._cache[name] = fn
        return fn
