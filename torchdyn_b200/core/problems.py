"""``ODEProblem``: binds a vector field to a solver and a sensitivity algorithm.

Reference: torchdyn/core/problems.py:18-159.  Kept: constructor signature and defaults (atol=rtol=1e-4,
adjoint 1e-6), ``solver_adjoint`` given BY NAME, the flat ``vf_params`` re-built at every forward so that
gradients flow into the individual ``nn.Parameter``s, the positional call of the autograd Function
(``(vf_params, x, t_span, save_at, args)`` -- the user's ``save_at`` lands in the unused ``B`` slot, as in the
reference), ``seminorm`` accepted but not wired (backsolve always uses it, interpolated never).
MultipleShootingProblem / SDEProblem are outside this implementation.
"""
from __future__ import annotations

import weakref
from typing import Callable, Generator, Iterable, Union

import torch
import torch.nn as nn
from torch import Tensor

from ..numerics.odeint import odeint
from ..numerics.sensitivity import _gather_odefunc_adjoint, _gather_odefunc_interp_adjoint
from ..numerics.solvers.ode import str_to_solver
from .utils import standardize_vf_call_signature

_SENSITIVITIES = ("autograd", "adjoint", "interpolated_adjoint")
# ODEProblem -> (configuration key, autograd Function class): building the class costs ~0.1 ms per forward otherwise; kept
# off the module so that it stays picklable
_FN_CACHE = weakref.WeakKeyDictionary()


class ODEProblem(nn.Module):
    def __init__(self, vector_field: Union[Callable, nn.Module], solver: Union[str, nn.Module],
                 interpolator: Union[str, Callable, None] = None, order: int = 1, atol: float = 1e-4,
                 rtol: float = 1e-4, sensitivity: str = "autograd",
                 solver_adjoint: Union[str, nn.Module, None] = None, atol_adjoint: float = 1e-6,
                 rtol_adjoint: float = 1e-6, seminorm: bool = False,
                 integral_loss: Union[Callable, None] = None,
                 optimizable_params: Union[Iterable, Generator] = ()):
        """An ODE problem: vector field + solver + sensitivity algorithm
        ('autograd' | 'adjoint' | 'interpolated_adjoint')."""
        super().__init__()
        if type(solver) == str:
            solver = str_to_solver(solver)          # tableau built in fp32 here, only CAST later (Appendix A.1)
        solver_adjoint = solver if solver_adjoint is None else str_to_solver(solver_adjoint)

        self.solver, self.interpolator, self.atol, self.rtol = solver, interpolator, atol, rtol
        self.solver_adjoint, self.atol_adjoint, self.rtol_adjoint = solver_adjoint, atol_adjoint, rtol_adjoint
        self.sensitivity, self.integral_loss = sensitivity, integral_loss

        vector_field = standardize_vf_call_signature(vector_field)
        self.vf, self.order, self.sensalg = vector_field, order, sensitivity
        optimizable_params = tuple(optimizable_params)

        if len(tuple(self.vf.parameters())) > 0:
            pass
        elif len(optimizable_params) > 0:
            for k, p in enumerate(optimizable_params):
                self.vf.register_parameter(f"optimizable_parameter_{k}", p)
        else:
            print("Your vector field does not have `nn.Parameters` to optimize.")
            self.vf.register_parameter("dummy_parameter", nn.Parameter(torch.zeros(1)))
        self.vf_params = self._flatten_params()

    def _flatten_params(self) -> Tensor:
        return torch.cat([p.contiguous().flatten() for p in self.vf.parameters()])

    def _autograd_func(self):
        "autograd.Function (its .apply) for the selected adjoint"
        self.vf_params = self._flatten_params()
        if self.sensalg == "adjoint":
            gather = _gather_odefunc_adjoint
        elif self.sensalg == "interpolated_adjoint":
            gather = _gather_odefunc_interp_adjoint
        else:
            return None
        # (the Function closes over the configuration, not over vf_params: re-used until an attribute is re-assigned)
        key = (gather, id(self.vf), id(self.solver), self.atol, self.rtol, id(self.interpolator) if not isinstance(self.interpolator, str) else self.interpolator,
               id(self.solver_adjoint), self.atol_adjoint, self.rtol_adjoint, id(self.integral_loss))
        hit = _FN_CACHE.get(self)
        if hit is None or hit[0] != key:
            hit = (key, gather(self.vf, self.vf_params, self.solver, self.atol, self.rtol, self.interpolator,
                               self.solver_adjoint, self.atol_adjoint, self.rtol_adjoint, self.integral_loss,
                               problem_type="standard"))
            _FN_CACHE[self] = hit
        return hit[1].apply

    def odeint(self, x: Tensor, t_span: Tensor, save_at: Tensor = (), args={}):
        "Returns Tuple(`t_eval`, `solution`)"
        if self.sensalg == "autograd":
            if isinstance(t_span, Tensor) and t_span.requires_grad and torch.is_grad_enabled():
                from warnings import warn
                warn("torchdyn_b200: gradients with respect to `t_span` are not propagated under sensitivity='autograd' "
                     "(step sizes are constants of the graph); use sensitivity='adjoint' for dL/dt_span")
            return odeint(self.vf, x, t_span, self.solver, self.atol, self.rtol, interpolator=self.interpolator,
                          save_at=save_at, args=args)
        fn = self._autograd_func()
        if fn is None:
            raise ValueError(f"unknown sensitivity {self.sensalg!r}; expected one of {_SENSITIVITIES}")
        return fn(self.vf_params, x, t_span, save_at, args)

    def forward(self, x: Tensor, t_span: Tensor, save_at: Tensor = (), args={}):
        "For safety redirects to intended method `odeint`"
        return self.odeint(x, t_span, save_at, args)
