"""Vector-field call-signature adapters + NFE counter.

Reference: torchdyn/core/defunc.py:19-35 (DEFuncBase), :38-94 (DEFunc).  ``nfe`` advances by one per
vector-field evaluation -- it is the unit of the throughput metric (sample-NFE/s) -- including when the
evaluation happens inside a fused CUDA kernel (the fused paths add their counts explicitly).
"""
from __future__ import annotations

from typing import Callable, Dict

import torch.nn as nn
from torch import Tensor, cat


class DEFuncBase(nn.Module):
    def __init__(self, vector_field: Callable, has_time_arg: bool = True):
        """Wraps a callable so that it is invoked as ``vf(t, x, args=...)`` (or ``vf(x)`` when it takes no time)."""
        super().__init__()
        self.nfe, self.vf, self.has_time_arg = 0.0, vector_field, has_time_arg

    def forward(self, t: Tensor, x: Tensor, args: Dict = {}) -> Tensor:
        self.nfe += 1
        return self.vf(t, x, args=args) if self.has_time_arg else self.vf(x)


class DEFunc(nn.Module):
    def __init__(self, vector_field: Callable, order: int = 1):
        """NeuralODE's vector-field wrapper: depth (``t``) injection into modules exposing a ``t`` attribute,
        integral-loss propagation in ``x[:, 0]`` for sensitivity='autograd', higher-order systems."""
        super().__init__()
        self.vf, self.nfe = vector_field, 0.0
        self.order, self.integral_loss, self.sensitivity = order, None, None
        self._depth_modules = None

    def _inject_depth(self, t):
        # the reference re-scans named_modules() on every evaluation (defunc.py:63-65); the set of modules
        # with a `t` attribute is fixed after construction, so scan once.
        if self._depth_modules is None:
            object.__setattr__(self, "_depth_modules", [m for _, m in self.vf.named_modules() if hasattr(m, "t")])
        for m in self._depth_modules:
            m.t = t

    def forward(self, t: Tensor, x: Tensor, args: Dict = {}) -> Tensor:
        self.nfe += 1
        self._inject_depth(t)
        if (self.integral_loss is not None) and self.sensitivity == "autograd":
            x_dyn = x[:, 1:]
            dlds = self.integral_loss(t, x_dyn)
            if len(dlds.shape) == 1:
                dlds = dlds[:, None]
            x_dyn = self.higher_order_forward(t, x_dyn) if self.order > 1 else self.vf(t, x_dyn)
            return cat([dlds, x_dyn], 1).to(x_dyn)
        if self.order > 1:
            return self.higher_order_forward(t, x)
        return self.vf(t, x, args=args)

    def higher_order_forward(self, t: Tensor, x: Tensor, args: Dict = {}) -> Tensor:
        width = x.size(1) // self.order
        parts = [x[:, width * i: width * (i + 1)] for i in range(1, self.order)]
        parts.append(self.vf(t, x))
        return cat(parts, dim=1).to(x)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_depth_modules"] = None
        return d
