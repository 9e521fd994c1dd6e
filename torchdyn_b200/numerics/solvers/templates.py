"""Solver plug-in contract (reference: torchdyn/numerics/solvers/templates.py:12-43, ode.py:27-58).

A solver is an ``nn.Module`` with
  * ``order``, ``stepping_class`` in {'fixed', 'adaptive'}, ``tableau = (c, [a_i], bsol, berr) | None``,
    ``safety / min_factor / max_factor`` one-element tensors (``max_factor`` is int64, like the reference's
    ``torch.tensor([10])``);
  * ``sync_device_dtype(x, t_span) -> (x, t_span)``;
  * ``step(f, x, t, dt, k1=None, args=None)`` -> adaptive ``(f_new, x_new, x_err, stages)``, fixed ``(None, x_new, None)``.
User subclasses that override ``step`` are run unchanged by ``odeint`` (torch ops, their own code); the four
built-ins (exact types) are executed by the CUDA kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class DiffEqSolver(nn.Module):
    def __init__(self, order, stepping_class: str = "fixed", min_factor: float = 0.2, max_factor: float = 10,
                 safety: float = 0.9):
        super().__init__()
        self.order = order
        self.stepping_class = stepping_class
        self.min_factor = torch.tensor([min_factor])
        self.max_factor = torch.tensor([max_factor])
        self.safety = torch.tensor([safety])
        self.tableau = None

    def sync_device_dtype(self, x, t_span):
        """Put the tableau on the state's device/dtype (a CAST of the stored tableau, not a rebuild --
        templates.py:30-40) and ``t_span`` on its device.  Controller constants stay where they are: the
        step-size controller of this implementation runs on the host in IEEE fp32."""
        if isinstance(x, dict):
            proto = x[next(iter(x))]
        elif isinstance(x, torch.Tensor):
            proto = x
        else:
            raise NotImplementedError(f"{type(x)} is not supported as the state variable")
        if self.tableau is not None:
            c, a, bsol, berr = self.tableau
            self.tableau = (c.to(proto), [r.to(proto) for r in a], bsol.to(proto), berr.to(proto))
        if isinstance(t_span, torch.Tensor):
            t_span = t_span.to(proto.device)
        return x, t_span

    def controller_constants(self):
        """(safety, min_factor, max_factor) as python floats."""
        return float(self.safety), float(self.min_factor), float(self.max_factor)

    def step(self, f, x, t, dt, k1=None, args=None):
        raise NotImplementedError("Stepping rule not implemented for the solver")
