"""reference: ps/src/ml/util/math_util.{hpp,cpp} (SafeLog, Sigmoid, LogSum, LogSumVec, Softmax, feature dot products,
FeatureScaleAndAdd)."""
from __future__ import annotations

import math
from typing import Sequence

import torch

from .features import DenseFeature, SparseFeature

_CUTOFF = 1e-10


def safe_log(x: float) -> float:
    """log with the argument clamped away from zero (log(1e-10) for anything smaller)."""
    return math.log(max(float(x), _CUTOFF))


def sigmoid(x: float) -> float:
    x = float(x)
    if x >= 0:
        return 1.0 / (1.0 + math.exp(-x))
    e = math.exp(x)
    return e / (1.0 + e)


def log_sum(log_a: float, log_b: float) -> float:
    """log(exp(a) + exp(b)) without overflow."""
    hi, lo = (log_a, log_b) if log_a > log_b else (log_b, log_a)
    return hi + math.log1p(math.exp(lo - hi))


def log_sum_vec(logvec: Sequence[float]) -> float:
    return float(torch.logsumexp(torch.as_tensor(logvec, dtype=torch.float64), 0))


def softmax(vec) -> torch.Tensor:
    """In the reference this normalises a vector of log-weights in place; here it returns the result."""
    return torch.softmax(torch.as_tensor(vec, dtype=torch.float32), 0)


def dot(f1, f2) -> float:
    """Dense·dense, dense·sparse, sparse·dense and sparse·sparse (merge join) products."""
    if isinstance(f1, DenseFeature) and isinstance(f2, DenseFeature):
        return float(f1.v @ f2.v)
    if isinstance(f1, SparseFeature) and isinstance(f2, DenseFeature):
        return float((f2.v[f1.ids] * f1.vals).sum())
    if isinstance(f1, DenseFeature) and isinstance(f2, SparseFeature):
        return dot(f2, f1)
    k = torch.searchsorted(f2.ids, f1.ids).clamp_max(max(f2.ids.numel() - 1, 0))
    if f2.ids.numel() == 0:
        return 0.0
    hit = f2.ids[k] == f1.ids
    return float((f1.vals[hit] * f2.vals[k[hit]]).sum())


def feature_scale_and_add(alpha: float, f1, f2: DenseFeature) -> None:
    """f2 += alpha * f1 (f1 dense or sparse)."""
    if isinstance(f1, DenseFeature):
        f2.v.add_(f1.v, alpha=alpha)
    else:
        f2.v.index_add_(0, f1.ids, f1.vals * alpha)
