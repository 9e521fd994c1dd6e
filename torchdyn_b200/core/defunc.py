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
        # the reference re-scans named_modules() on every evaluation (defunc.py:63-65); here the list of modules with a
        # `t` attribute is kept and re-derived whenever the module tree changed (a swapped or added layer)
        mods = tuple(self.vf.modules()) if isinstance(self.vf, nn.Module) else ()
        key = tuple(map(id, mods))
        if self._depth_modules is None or self._depth_modules[0] != key:
            object.__setattr__(self, "_depth_modules", (key, [m for m in mods if hasattr(m, "t")]))
        for m in self._depth_modules[1]:
            m.t = t

    def forward(self, t: Tensor, x: Tensor, args: Dict = {}) -> Tensor:
        self.nfe += 1
        self._inject_depth(t)
        carries_cost = self.integral_loss is not None and self.sensitivity == "autograd"
        if not carries_cost:
            return self.higher_order_forward(t, x) if self.order > 1 else self.vf(t, x, args=args)
        # back-propagation through the solver with an integral loss: column 0 of the state accumulates the running cost
        # g(t, z), the remaining columns z follow the dynamics (reference core/defunc.py:68-77)
        z = x[:, 1:]
        running = self.integral_loss(t, z)
        dz = self.higher_order_forward(t, z) if self.order > 1 else self.vf(t, z)
        return cat([running[:, None] if running.dim() == 1 else running, dz], 1).to(dz)

    def higher_order_forward(self, t: Tensor, x: Tensor, args: Dict = {}) -> Tensor:
        width = x.size(1) // self.order
        parts = [x[:, width * i: width * (i + 1)] for i in range(1, self.order)]
        parts.append(self.vf(t, x))
        return cat(parts, dim=1).to(x)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_depth_modules"] = None
        return d
