"""Learning-rate policies.

reference: src/caffe/solver.cpp:767-790 (GetLearningRate):
  fixed: base_lr                       step: base_lr * gamma^floor(iter/stepsize)
  exp:   base_lr * gamma^iter          inv:  base_lr * (1 + gamma*iter)^(-power)
  poly:  base_lr * (1 - iter/max_iter)^power
"""
from __future__ import annotations

import math


def learning_rate(sp, it: int) -> float:
    policy = sp.lr_policy
    base = float(sp.base_lr)
    if policy == "fixed":
        return base
    if policy == "step":
        if not sp.stepsize or sp.stepsize <= 0:
            raise ValueError("lr_policy 'step' requires a positive stepsize")
        return base * math.pow(sp.gamma, it // sp.stepsize)
    if policy == "exp":
        return base * math.pow(sp.gamma, it)
    if policy == "inv":
        return base * math.pow(1.0 + sp.gamma * it, -sp.power)
    if policy == "poly":
        return base * math.pow(1.0 - float(it) / float(sp.max_iter), sp.power)
    raise ValueError(f"Unknown learning rate policy: {policy}")
