"""The four hot-path steppers -- euler, rk4, dopri5, tsit5 -- executed by CUDA kernels.

Reference: torchdyn/numerics/solvers/ode.py:61-71 (Euler), :89-104 (RungeKutta4), :138-156
(DormandPrince45), :160-178 (Tsitouras45), registry :320-336.  Each stepper is described by a *plan*: for
every stage the combination shape the reference expression has (CHAIN: ``x + dt*a0*k0 + dt*a1*k1 ...``
summed left to right; INNER: ``x + dt*(a0*k0 + a1*k1 ...)``), because PyTorch rounds each op separately and
the accepted-step sequence depends on those bits.  ``tdyn_rk_combine`` / ``tdyn_rk_finish`` reproduce the
roundings exactly.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Tuple

import numpy as np
import torch

from ... import _kernels
from ..._lib import MODE_CHAIN, MODE_INNER
from ._constants import construct_dopri5, construct_rk4, construct_tsit5
from .templates import DiffEqSolver

f32 = np.float32


class Stage(NamedTuple):
    mode: int
    coefs: Tuple[float, ...]
    c: "np.float32"            # t_stage = t + c*dt


class Plan(NamedTuple):
    stages: Tuple[Stage, ...]
    bsol: Tuple[float, ...]
    berr: Optional[Tuple[float, ...]]
    fsal: bool


def _floats(t: torch.Tensor) -> Tuple[float, ...]:
    return tuple(t.detach().to("cpu", torch.float32).tolist())


class Reversed:
    """f_(t, x) = -f(-t, x): the vector field of a reversed-time solve (reference numerics/odeint.py:54-58).
    Library-internal vector fields that accept host-side times (``tdyn_eval``) keep doing so through it."""

    def __init__(self, f):
        self.f = f
        inner = getattr(f, "tdyn_eval", None)
        if inner is not None:
            import inspect
            if "sign" in inspect.signature(inner).parameters:     # kernels that fold the negation into their epilogue
                self.tdyn_eval = lambda t_host, y: inner(-t_host, y, sign=-1.0)
            else:
                self.tdyn_eval = lambda t_host, y: -inner(-t_host, y)

    def __call__(self, t, x):
        return -self.f(-t, x)


def call_vf(kern, f, t_host, y):
    """Evaluate the vector field at host-side fp32 time ``t_host`` and make the result kernel-ready.
    Library-internal vector fields expose ``tdyn_eval(t_host, y)`` (no device time tensor, no sync); any
    other callable gets the shape-[1] device tensor the reference would pass.  When autograd is recording
    (sensitivity='autograd'), the result keeps its graph: the stage combinations are differentiable
    (``rk_combine`` below)."""
    h = getattr(f, "tdyn_eval", None)
    k = h(t_host, y) if h is not None else f(kern.time_tensor(t_host, y), y)
    if k.dtype != y.dtype:
        k = k.to(y.dtype)
    return k if k.is_contiguous() else k.contiguous()


class _CombineFn(torch.autograd.Function):
    """y = combine(x, ks, coefs, mode, dt) on the CUDA kernel, differentiable w.r.t. x and every k_j
    (sensitivity='autograd', reference core/problems.py:142-153: back-propagation THROUGH the solver).  The map is
    linear: dy/dx = 1, dy/dk_j = dt*coef_j; dt is a constant of the graph (the reference computes its step sizes
    under no_grad, numerics/utils.py:57, except for init_step's dependence of dt0 on x0, which is dropped here)."""

    @staticmethod
    def forward(ctx, kern, coefs, mode, dt, x, *ks):
        ctx.scales = [float(f32(f32(dt) * f32(c))) for c in coefs]
        ctx.has_x = x is not None
        y = kern.combine(kern.empty_like(ks[0] if x is None else x), None if x is None else x.detach(),
                         [k.detach() if k.is_contiguous() else k.detach().contiguous() for k in ks], coefs, mode, dt)
        return y

    @staticmethod
    def backward(ctx, g):
        gx = g if ctx.has_x else None
        return (None, None, None, None, gx) + tuple(g * s for s in ctx.scales)


def rk_combine(kern, x, ks, coefs, mode, dt):
    """Stage combination: plain kernel launch, or its differentiable wrapper when autograd is recording."""
    if torch.is_grad_enabled() and ((x is not None and x.requires_grad) or any(k.requires_grad for k in ks)):
        return _CombineFn.apply(kern, tuple(coefs), mode, float(dt), x, *ks)
    like = ks[0] if x is None else x
    return kern.combine(kern.empty_like(like), x, ks, coefs, mode, dt)


class _KernelSolver(DiffEqSolver):
    """Shared machinery of the built-in steppers."""

    def __init__(self, order, stepping_class, dtype):
        super().__init__(order=order, stepping_class=stepping_class)
        self.dtype = dtype
        self._plan_cache = None

    def _build_plan(self) -> Plan:
        raise NotImplementedError

    def plan(self) -> Plan:
        """fp32 coefficient plan of the CURRENT tableau tensors (they may have been cast by sync_device_dtype)."""
        key = None if self.tableau is None else tuple((id(t), t.dtype) for t in (self.tableau[0], self.tableau[2]))
        if self._plan_cache is None or self._plan_cache[0] != key:
            self._plan_cache = (key, self._build_plan())
        return self._plan_cache[1]

    def __setstate__(self, state):
        super().__setstate__(state)
        self._plan_cache = None   # keyed on tensor identities, which do not survive pickling

    # -- kernels driven from the host -------------------------------------------------------------
    def stages(self, kern, f, x, t, dt, k1) -> List[torch.Tensor]:
        """k1..ks for one attempted step at host-side fp32 ``t``/``dt``."""
        p = self.plan()
        ks = [k1 if k1 is not None else call_vf(kern, f, t, x)]
        for st in p.stages:
            y = rk_combine(kern, x, ks[:len(st.coefs)], st.coefs, st.mode, dt)
            ks.append(call_vf(kern, f, f32(t + st.c * dt), y))
        return ks

    @staticmethod
    def _host_scalar(v) -> "np.float32":
        return f32(v.item() if isinstance(v, torch.Tensor) else v)

    def step(self, f, x, t, dt, k1=None, args=None):
        kern = _kernels.get(x)
        t_h, dt_h = self._host_scalar(t), self._host_scalar(dt)
        x = x if x.is_contiguous() else x.contiguous()
        p = self.plan()
        ks = self.stages(kern, f, x, t_h, dt_h, k1)
        x_new = rk_combine(kern, x, ks[:len(p.bsol)], p.bsol, MODE_INNER if len(p.bsol) > 1 else MODE_CHAIN, dt_h)
        if p.berr is None:
            return None, x_new, None
        err = rk_combine(kern, None, ks[:len(p.berr)], p.berr, MODE_INNER, dt_h)
        return ks[-1], x_new, err, tuple(ks)


class Euler(_KernelSolver):
    def __init__(self, dtype=torch.float32):
        """Explicit Euler, order 1 (ode.py:61-71): x + dt*k1."""
        super().__init__(1, "fixed", dtype)

    def _build_plan(self):
        return Plan((), (1.0,), None, False)


class Midpoint(_KernelSolver):
    def __init__(self, dtype=torch.float32):
        """Explicit midpoint, order 2 (ode.py:75-86): x_mid = x + 0.5*dt*k1; x + dt*f(t + 0.5*dt, x_mid)."""
        super().__init__(2, "fixed", dtype)

    def _build_plan(self):
        # x + dt*(0*k1 + 1*k2) has the bits of x + dt*k2
        return Plan((Stage(MODE_CHAIN, (0.5,), f32(0.5)),), (0.0, 1.0), None, False)


class RungeKutta4(_KernelSolver):
    def __init__(self, dtype=torch.float32):
        """Classic RK4 (ode.py:89-104); every stage is x + dt*(a0*k1 + ...) including the literal zeros."""
        super().__init__(4, "fixed", dtype)
        self.tableau = construct_rk4(dtype)

    def _build_plan(self):
        c, a, bsol, _ = self.tableau
        cs = _floats(c)
        stages = tuple(Stage(MODE_INNER, _floats(a[i]), f32(cs[i])) for i in range(3))
        return Plan(stages, _floats(bsol), None, False)


class _Embedded54(_KernelSolver):
    """7-stage FSAL 5(4) pairs (ode.py:145-156, :167-178): stage 2 and 4..7 are CHAIN, stage 3 is INNER,
    x_new = x + dt*(sum_{j<6} bsol_j k_j), err = dt*(sum_{j<7} berr_j k_j)."""

    def _build_plan(self):
        c, a, bsol, berr = self.tableau
        cs = _floats(c)
        stages = []
        for i in range(6):
            stages.append(Stage(MODE_INNER if i == 1 else MODE_CHAIN, _floats(a[i]), f32(cs[i])))
        return Plan(tuple(stages), _floats(bsol)[:6], _floats(berr), True)


class DormandPrince45(_Embedded54):
    def __init__(self, dtype=torch.float32):
        super().__init__(5, "adaptive", dtype)
        self.tableau = construct_dopri5(dtype)


class Tsitouras45(_Embedded54):
    def __init__(self, dtype=torch.float32):
        super().__init__(5, "adaptive", dtype)
        self.tableau = construct_tsit5(dtype)


class AsynchronousLeapfrog(DiffEqSolver):
    """Explicit asynchronous-leapfrog stepper on a phase-space state ``[x, v]`` with ``v' = f(t, x)`` (reference
    ode.py:107-135).  A plug-in solver: its ``step`` is torch code that ``odeint`` / ``odeint_symplectic`` drive through
    the solver contract (f stays a torch call).  Only ``stepping_class='fixed'`` integrates: the reference's adaptive
    variant returns a 3-tuple where its adaptive loop unpacks four values, i.e. it cannot run there either."""

    def __init__(self, channel_index: int = -1, stepping_class: str = 'fixed', dtype=torch.float32):
        super().__init__(order=2)
        self.dtype = dtype
        self.channel_index = channel_index
        self.stepping_class = stepping_class
        self.const = 1
        self.tableau = construct_rk4(self.dtype)
        self.x_shape = None

    def step(self, f, xv, t, dt, k1=None, args=None):
        half = xv.shape[-1] // 2
        x, v = xv[..., :half], xv[..., half:]
        if k1 is None:
            k1 = f(t, x)                      # evaluated and not used, as in the reference (it is an NFE there too)
        x_mid = x + 0.5 * dt * v
        v_mid = f(t + 0.5 * dt, x_mid)
        v_new = 2 * self.const * (v_mid - v) + v
        x_new = x_mid + 0.5 * dt * v_new
        xv_err = torch.cat([torch.zeros_like(x), v], -1) if self.stepping_class == 'adaptive' else None
        return None, torch.cat([x_new, v_new], -1), xv_err


BUILTIN_TYPES = (Euler, Midpoint, RungeKutta4, DormandPrince45, Tsitouras45)

# name registry of the explicit solvers (reference ode.py:320-325; ieuler and the multiple-shooting solvers are outside
# this implementation: unknown names raise KeyError as there)
SOLVER_DICT = {'euler': Euler, 'midpoint': Midpoint, 'alf': AsynchronousLeapfrog, 'AsynchronousLeapfrog': AsynchronousLeapfrog,
               'rk4': RungeKutta4, 'rk-4': RungeKutta4, 'RungeKutta4': RungeKutta4,
               'dopri5': DormandPrince45, 'DormandPrince45': DormandPrince45, 'DormandPrince5': DormandPrince45,
               'tsit5': Tsitouras45, 'Tsitouras45': Tsitouras45, 'Tsitouras5': Tsitouras45}


def str_to_solver(solver_name, dtype=torch.float32):
    "Solver name -> solver instance with its tableau built in `dtype`."
    return SOLVER_DICT[solver_name](dtype)
