"""``NeuralODE``: the user-facing module (reference torchdyn/core/neuralde.py:29-152).

``NeuralODE(vector_field, solver='tsit5', order=1, atol=1e-3, rtol=1e-3, sensitivity='autograd',
solver_adjoint=None, atol_adjoint=1e-4, rtol_adjoint=1e-4, interpolator=None, integral_loss=None,
seminorm=False, return_t_eval=True, optimizable_params=())``; ``forward(x, t_span=None, save_at=(), args={})``
returns ``(t_eval, sol)`` (or ``sol`` when ``return_t_eval=False``); ``trajectory(x, t_span)`` returns ``sol``.
Derives from ``pytorch_lightning.LightningModule`` too when that package is importable (optional sugar).
"""
from __future__ import annotations

from typing import Callable, Dict, Generator, Iterable, Union

import torch
import torch.nn as nn
from torch import Tensor

from ..numerics.odeint import odeint
from .problems import ODEProblem
from .utils import standardize_vf_call_signature

try:  # optional
    import pytorch_lightning as _pl
    _Bases = (ODEProblem, _pl.LightningModule)
except Exception:  # pragma: no cover - not installed in this image
    _Bases = (ODEProblem,)


class NeuralODE(*_Bases):
    def __init__(self, vector_field: Union[Callable, nn.Module], solver: Union[str, nn.Module] = "tsit5",
                 order: int = 1, atol: float = 1e-3, rtol: float = 1e-3, sensitivity="autograd",
                 solver_adjoint: Union[str, nn.Module, None] = None, atol_adjoint: float = 1e-4,
                 rtol_adjoint: float = 1e-4, interpolator: Union[str, Callable, None] = None,
                 integral_loss: Union[Callable, None] = None, seminorm: bool = False, return_t_eval: bool = True,
                 optimizable_params: Union[Iterable, Generator] = ()):
        super().__init__(vector_field=standardize_vf_call_signature(vector_field, order, defunc_wrap=True),
                         order=order, sensitivity=sensitivity, solver=solver, atol=atol, rtol=rtol,
                         solver_adjoint=solver_adjoint, atol_adjoint=atol_adjoint, rtol_adjoint=rtol_adjoint,
                         seminorm=seminorm, interpolator=interpolator, integral_loss=integral_loss,
                         optimizable_params=optimizable_params)
        self._control, self.controlled, self.t_span = None, False, None
        self.return_t_eval = return_t_eval
        if integral_loss is not None:
            self.vf.integral_loss = integral_loss
        self.vf.sensitivity = sensitivity

    def _prep_integration(self, x: Tensor, t_span: Tensor):
        "Default t_span, Hutchinson noise sampling (once per forward) and data-control assignment."
        if t_span is None:
            t_span = torch.linspace(0, 1, 2) if self.t_span is None else self.t_span

        excess_dims = 0
        if (self.integral_loss is not None) and self.sensitivity == "autograd":
            excess_dims += 1
        for _, module in self.vf.named_modules():
            if hasattr(module, "trace_estimator"):
                if module.noise_dist is not None:
                    module.noise = module.noise_dist.sample((x.shape[0],))
                excess_dims += 1
            if hasattr(module, "_control"):
                self.controlled = True
                module._control = x[:, excess_dims:].detach()
        return x, t_span

    def forward(self, x: Union[Tensor, Dict], t_span: Tensor = None, save_at: Iterable = (), args={}):
        x, t_span = self._prep_integration(x, t_span)
        t_eval, sol = super().forward(x, t_span, save_at, args)
        return (t_eval, sol) if self.return_t_eval else sol

    def trajectory(self, x: torch.Tensor, t_span: Tensor):
        x, t_span = self._prep_integration(x, t_span)
        _, sol = odeint(self.vf, x, t_span, solver=self.solver, atol=self.atol, rtol=self.rtol)
        return sol

    def __repr__(self):
        npar = sum([p.numel() for p in self.vf.parameters()])
        return (f"Neural ODE:\n\t- order: {self.order}"
                f"\n\t- solver: {self.solver}\n\t- adjoint solver: {self.solver_adjoint}"
                f"\n\t- tolerances: relative {self.rtol} absolute {self.atol}"
                f"\n\t- adjoint tolerances: relative {self.rtol_adjoint} absolute {self.atol_adjoint}"
                f"\n\t- num_parameters: {npar}"
                f"\n\t- NFE: {self.vf.nfe}")
