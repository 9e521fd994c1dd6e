"""Functional ODE integration API: ``odeint(f, x, t_span, solver, atol, rtol, ...) -> (t_eval, sol)``.

Drop-in for torchdyn.numerics.odeint (reference numerics/odeint.py:29-91 dispatch, :326-408 adaptive loop,
:411-445 fixed loop) for the solvers euler / rk4 / dopri5 / tsit5.

How it runs on a B200
  * the state, every RK stage and all outputs stay in HBM; each stage input is ONE fused kernel
    (``tdyn_rk_combine``), the end of a step is ONE kernel (``tdyn_rk_finish``: x_new, embedded error, scaled
    squared-error reduction) -- 168 B per state element per dopri5/tsit5 step instead of the ~770 B the
    reference's ~78 element-wise launches move;
  * the step-size controller (t, dt, checkpoint bookkeeping, accept/reject) runs on the host in IEEE fp32,
    fed by ONE 8-byte device->host read per attempted step (the reference syncs ~7 times per step);
  * MLP vector fields (``nn.Sequential`` of Linear + Tanh/Softplus/ReLU) are recognised and handed to the
    fused kernels (``torchdyn_b200.fused``): the vector field itself then also runs in this library.
Semantics that change the accepted-step sequence are reproduced on purpose:
per-op fp32 rounding of the stage expressions, ``dt_old - dt`` after a checkpoint clamp, exact float
equality for hitting output times, reversed-time handling (negated span, un-negated f in init_step).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Tuple, Union
from warnings import warn

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .. import _kernels
from .. import distributed as _dist
from .interpolators import FourthOrder, str_to_interp
from .solvers.ode import BUILTIN_TYPES, Reversed, Tsitouras45, call_vf, rk_combine, str_to_solver
from .utils import adapt_step, init_step, rms_from_sumsq

f32 = np.float32
MAX_STEPS = 1_000_000   # safety valve (the reference loops forever on a NaN error ratio)

# Optional per-attempt trace: set to a list and every adaptive solve appends
# {"dt0", "t": [...], "dt": [...], "ratio": [...], "accept": [...]} (used by the parity tests and bench.py).
TRACE_SINK = None


def _host_times(t_span) -> np.ndarray:
    if isinstance(t_span, (list, tuple)):
        t_span = torch.cat([torch.as_tensor(t).reshape(-1) for t in t_span])
    if isinstance(t_span, Tensor):
        if t_span.dim() > 1:
            raise NotImplementedError("Parallel routines not implemented yet, check experimental versions of `torchdyn`")
        return t_span.detach().to("cpu", torch.float32).numpy().copy()
    return np.asarray(t_span, dtype=np.float32)


def _times_tensor(times: List, like: Tensor) -> Tensor:
    return torch.tensor(np.asarray(times, dtype=np.float32), dtype=like.dtype, device=like.device)


def odeint(f: Callable, x: Tensor, t_span: Union[List, Tensor], solver: Union[str, nn.Module], atol: float = 1e-3,
           rtol: float = 1e-3, t_stops: Union[List, Tensor, None] = None, verbose: bool = False,
           interpolator: Union[str, Callable, None] = None, return_all_eval: bool = False,
           save_at: Union[Iterable, Tensor] = (), args: Dict = {},
           seminorm: Tuple[bool, Union[int, None]] = (False, None)) -> Tuple[Tensor, Tensor]:
    """Solve the IVP dx/dt = f(t, x), x(t_span[0]) = x.  Same arguments, warnings, errors and return
    convention as the reference (``(t_eval, sol)`` with ``sol[i]`` the state at ``t_eval[i]``)."""
    ts = _host_times(t_span)
    if isinstance(x, Tensor) and x.is_cuda and (type(solver) == str or type(solver) in BUILTIN_TYPES) \
            and not torch.is_grad_enabled():
        from .. import fused
        if not hasattr(f, "tdyn_eval"):     # (the adjoints' own dynamics objects keep the label of the user-level solve)
            fused.reset_last()
        f = fused.wrap(f, x)            # MLP vector fields are evaluated by this library's kernels (no autograd graph)
    if ts[1] < ts[0]:   # reversed time: integrate -f(-t, x) over the negated span (odeint.py:54-58)
        if verbose:
            warn("You are integrating on a reversed time domain, adjusting the vector field automatically")
        f_ = Reversed(f)
        ts = -ts
    else:
        f_ = f

    if type(solver) == str:
        solver = str_to_solver(solver, x.dtype if isinstance(x, Tensor) else torch.float32)
    x, _ = solver.sync_device_dtype(x, None)
    stepping_class = solver.stepping_class

    if isinstance(solver, Tsitouras45):
        if verbose:
            warn("Running interpolation not yet implemented for `tsit5`")
        interpolator = None
    if type(interpolator) == str:
        interpolator = str_to_interp(interpolator, x.dtype)
        x, _ = interpolator.sync_device_dtype(x, None)

    if stepping_class == 'fixed':
        if atol != odeint.__defaults__[0] or rtol != odeint.__defaults__[1]:
            warn("Setting tolerances has no effect on fixed-step methods")
        return _fixed_odeint(f_, x, ts, solver, save_at=save_at, args=args)
    elif stepping_class == 'adaptive':
        if isinstance(x, Tensor) and not x.is_contiguous():
            x = x.contiguous()
        t0 = f32(ts[0])
        if type(solver) in BUILTIN_TYPES and isinstance(x, Tensor) and x.is_cuda:
            from .. import fused
            done = fused.try_adaptive(f_, x, ts, solver, atol, rtol, interpolator, return_all_eval, seminorm)
            if done is not None:
                if len(save_at) > 0:
                    warn("Setting save_at has no effect on adaptive-step methods")
                return done
        if type(solver) in BUILTIN_TYPES:
            kern = _kernels.get(x)
            k1 = call_vf(kern, f_, t0, x)
            dt = init_step(f, k1, x, t0, solver.order, atol, rtol)      # un-negated f: reference quirk
            from . import graph_loop
            if graph_loop.ENABLED and isinstance(x, Tensor) and x.is_cuda:
                # opaque f: the whole accept / reject loop as one CUDA-graph launch (no host sync per step)
                done = graph_loop.try_device_loop(f, f_, k1, x, dt, ts, solver, atol, rtol, interpolator, return_all_eval, seminorm)
                if done is not None:
                    if len(save_at) > 0:
                        warn("Setting save_at has no effect on adaptive-step methods")
                    return done
        else:
            t0_t = torch.full((), float(t0), dtype=x.dtype, device=x.device)
            k1 = f_(t0_t, x)
            dt = _init_step_torch(f, k1, x, t0_t, solver.order, atol, rtol)
        if len(save_at) > 0:
            warn("Setting save_at has no effect on adaptive-step methods")
        return _adaptive_odeint(f_, k1, x, dt, ts, solver, atol, rtol, args, interpolator, return_all_eval, seminorm)
    raise ValueError(f"unknown stepping_class {stepping_class!r}")


def odeint_symplectic(f: Callable, x: Tensor, t_span: Union[List, Tensor], solver: Union[str, nn.Module], atol: float = 1e-3,
                      rtol: float = 1e-3, verbose: bool = False, return_all_eval: bool = False,
                      save_at: Union[List, Tensor] = ()):
    """IVP on a phase-space state with a symplectic stepper (reference odeint.py:95-153): same checks and errors, then the
    fixed-step loop of ``odeint`` (the solver is a plug-in: its torch ``step`` runs unchanged).  The reference's adaptive
    branch cannot run (its ALF step returns three values where the adaptive loop unpacks four): it raises here."""
    from .solvers.ode import AsynchronousLeapfrog
    if type(solver) == str:
        solver = str_to_solver(solver, x.dtype)
    if not hasattr(f, 'order'):
        raise RuntimeError('The system order should be specified as an attribute `order` of `vector_field`')
    if isinstance(solver, AsynchronousLeapfrog) and f.order == 2:
        raise RuntimeError('ALF solver should be given a vector field specified as a first-order symplectic system: v = f(t, x)')
    solver.x_shape = x.shape[-1] // 2
    if len(_host_times(t_span).shape) > 1:
        raise NotImplementedError("Parallel routines not implemented yet, check experimental versions of `torchdyn`")
    if solver.stepping_class != 'fixed':
        raise NotImplementedError("adaptive symplectic stepping: the reference's adaptive ALF step is incompatible with its "
                                  "own adaptive loop; use stepping_class='fixed'")
    if atol != odeint_symplectic.__defaults__[0] or rtol != odeint_symplectic.__defaults__[1]:
        warn("Setting tolerances has no effect on fixed-step methods")
    return odeint(f, x, t_span, solver, verbose=verbose, return_all_eval=return_all_eval, save_at=save_at)


class EventState:
    """Flags of the hybrid-system callbacks at one point (reference numerics/utils.py:76-82)."""

    def __init__(self, evid):
        self.evid = [bool(e) for e in evid]

    def __ne__(self, other):
        return sum(a != b for a, b in zip(self.evid, other.evid))

    def newly_triggered(self, previous: "EventState") -> int:
        """how many flags went from False (previous) to True (self)"""
        return sum((a != b) and (b is False) for a, b in zip(self.evid, previous.evid))


def odeint_hybrid(f, x, t_span, j_span, solver, callbacks, atol=1e-3, rtol=1e-3, event_tol=1e-4, priority='jump',
                  seminorm: Tuple[bool, Union[int, None]] = (False, None)):
    """IVP with state jumps defined by callbacks (``check_event(t, x) -> bool``, ``jump_map(t, x) -> x``); reference
    odeint.py:183-323.  Same loop: adaptive steps of an embedded pair (stages, error norm and controller on this library's
    kernels / host fp32 scalars, exactly as ``odeint``), after every attempted step the event flags are evaluated at the
    step end, a False -> True transition is closed in on by bisection of the step (``event_tol``), the state is saved
    before and after the jump.  Quirks kept: the flags are compared with those of the INITIAL point throughout (the
    reference never updates ``event_states``), ``j_span`` only counts the initial jump, rejected steps still run the
    controller, every accepted step is returned (``t_span`` interior points are hit exactly by the checkpoint clamp)."""
    ts = _host_times(t_span)
    if type(solver) == str:
        solver = str_to_solver(solver, x.dtype)
    x, _ = solver.sync_device_dtype(x, None)
    x_shape = x.shape
    builtin = type(solver) in BUILTIN_TYPES
    t_eval, t, T = ts[1:], f32(ts[0]), f32(ts[-1])
    tt = lambda v: torch.full((1,), float(v), dtype=x.dtype, device=x.device)      # noqa: E731  (callbacks see t as a [1] tensor)
    ck, ck_flag, dt_keep, jnum = 0, False, None, 0

    events0 = EventState([False for _ in callbacks])
    if priority == 'jump':
        ev = EventState([cb.check_event(tt(t), x) for cb in callbacks])
        if ev.newly_triggered(events0) > 0:
            i = min(i for i, flag in enumerate(ev.evid) if flag)
            x = callbacks[i].jump_map(tt(t), x)
            jnum += 1

    if builtin:
        if isinstance(x, Tensor) and not x.is_contiguous():
            x = x.contiguous()
        kern = _kernels.get(x)
        k1 = call_vf(kern, f, t, x)
        dt = init_step(f, k1, x, t, solver.order, atol, rtol)
        attempt = _KernelAttempt(solver, x, atol, rtol, seminorm, None)
    else:
        k1 = f(tt(t)[0], x)
        dt = _init_step_torch(f, k1, x, tt(t)[0], solver.order, atol, rtol)
        attempt = _TorchAttempt(solver, x, atol, rtol, seminorm, None, None)
    safety, min_factor, max_factor = solver.controller_constants()

    times, sol = [t], [x]
    n_attempts = 0
    while t < T and jnum < j_span:
        n_attempts += 1
        if n_attempts > MAX_STEPS or np.isnan(dt):
            raise RuntimeError(f"torchdyn_b200.odeint_hybrid: loop did not terminate (t={t}, dt={dt})")
        if f32(t + dt) > T:
            dt = f32(T - t)
        if ck < len(t_eval) and f32(t + dt) > t_eval[ck]:
            dt_keep, ck_flag = dt, True
            dt = f32(t_eval[ck] - t)
            ck += 1

        f_new, x_new, _, ratio = attempt(f, x, t, dt, k1)
        ev = EventState([cb.check_event(tt(f32(t + dt)), x_new) for cb in callbacks])

        if ev.newly_triggered(events0) > 0:
            # close in on the switching state inside [t, t + dt] by bisection of the step
            dt_pre, t_in, dt_in, x_in, niters = dt, t, dt, x, 0
            while niters < 100 and event_tol < dt_in:
                with torch.no_grad():
                    dt_in = f32(dt_in / f32(2))
                    f_new, x_try, _, _ = attempt(f, x_in, t_in, dt_in, k1)
                    ev = EventState([cb.check_event(tt(f32(t_in + dt_in)), x_try) for cb in callbacks])
                    niters += 1
                if ev.newly_triggered(events0) == 0:      # no event yet: advance the start of the search
                    x_in, t_in, dt_in, k1 = x_try, f32(t_in + dt_in), dt, f_new
            x, t = x_in, t_in
            i = min(i for i, flag in enumerate(ev.evid) if flag)
            sol.append(x.reshape(x_shape)); times.append(t)          # state and time BEFORE the jump
            x = callbacks[i].jump_map(tt(t), x)
            sol.append(x.reshape(x_shape)); times.append(t)          # ... and AFTER it
            k1, dt = None, dt_pre                                    # the next step evaluates f at the jumped state
        else:
            if ratio <= 1:
                t, x, k1 = f32(t + dt), x_new, f_new
                sol.append(x.reshape(x_shape)); times.append(t)
            if ck_flag:
                dt = f32(dt_keep - dt)
                ck_flag = False
            dt = adapt_step(dt, ratio, safety, min_factor, max_factor, solver.order)
    return _times_tensor(times, x), torch.stack(sol)


# ----------------------------------------------------------------------------------------------
# adaptive loop
# ----------------------------------------------------------------------------------------------

def _seminorm_count(x: Tensor, seminorm) -> int:
    if not seminorm[0]:
        return x.numel()
    if x.dim() == 0:
        return 1
    rows = min(int(seminorm[1]), x.shape[0])
    return rows * (x.numel() // max(x.shape[0], 1))


class _KernelAttempt:
    """One attempted step of a built-in embedded pair on the CUDA kernels."""

    def __init__(self, solver, x, atol, rtol, seminorm, interpolator):
        self.solver, self.atol, self.rtol = solver, atol, rtol
        self.kern = _kernels.get(x)
        self.plan = solver.plan()
        self.norm_n = _dist.local_norm_count(_seminorm_count(x, seminorm))
        self.interp = interpolator
        self._bmid = None

    def __call__(self, f, x, t, dt, k1):
        kern, p = self.kern, self.plan
        ks = self.solver.stages(kern, f, x, t, dt, k1)
        x_new = kern.empty_like(x)
        dks = [k.detach() for k in ks]
        kern.finish(x_new, None, x.detach(), dks, p.bsol, p.berr, dt, self.atol, self.rtol, self.norm_n)
        (sumsq,), count = _dist.reduce_sums(kern, 1, self.norm_n)
        if torch.is_grad_enabled() and (x.requires_grad or any(k.requires_grad for k in ks)):
            # sensitivity='autograd': the new state must stay on the graph (same values as the fused finish)
            x_new = rk_combine(kern, x, ks[:len(p.bsol)], p.bsol, 1, dt)
        return ks[-1], x_new, ks, rms_from_sumsq(sumsq, count, kern)

    def dense(self, x, x_new, ks, t, dt, t_out):
        if type(self.interp) is FourthOrder and not (torch.is_grad_enabled() and (x_new.requires_grad or x.requires_grad)):
            if self._bmid is None:
                self._bmid = tuple(self.interp.bmid.detach().to("cpu", torch.float32).tolist())
            th = f32(f32(t_out - t) / f32(f32(t + dt) - t))
            th2 = f32(th * th); th3 = f32(th2 * th); th4 = f32(th3 * th)
            return self.kern.dense_eval(self.kern.empty_like(x), x, x_new, list(ks), self._bmid, dt, (th, th2, th3, th4))
        return _dense_torch(self.interp, x, x_new, ks, t, dt, t_out)


class _TorchAttempt:
    """One attempted step of a USER solver module: its own ``step`` (torch ops) and the reference error norm."""

    def __init__(self, solver, x, atol, rtol, seminorm, interpolator, args):
        self.solver, self.atol, self.rtol, self.seminorm, self.interp, self.args = solver, atol, rtol, seminorm, interpolator, args

    def __call__(self, f, x, t, dt, k1):
        tt = torch.full((1,), float(t), dtype=x.dtype, device=x.device)
        dtt = torch.full((1,), float(dt), dtype=x.dtype, device=x.device)
        f_new, x_new, err, ks = self.solver.step(f, x, tt, dtt, k1=k1, args=self.args)
        if self.seminorm[0]:
            m = self.seminorm[1]
            es = err[:m] / (self.atol + self.rtol * torch.max(x[:m].abs(), x_new[:m].abs()))
        else:
            es = err / (self.atol + self.rtol * torch.max(x.abs(), x_new.abs()))
        return f_new, x_new, ks, f32(es.abs().pow(2).mean().sqrt().item())

    def dense(self, x, x_new, ks, t, dt, t_out):
        return _dense_torch(self.interp, x, x_new, ks, t, dt, t_out)


def _dense_torch(interp, x, x_new, ks, t, dt, t_out):
    mk = lambda v: torch.full((1,), float(v), dtype=x.dtype, device=x.device)  # noqa: E731
    dtt = mk(dt)
    x_mid = x + dtt * sum([interp.bmid[i] * ks[i] for i in range(len(ks))])
    coefs = interp.fit(dtt, ks[0], ks[-1], x, x_new, x_mid)
    return interp.evaluate(coefs, mk(t), mk(f32(t + dt)), mk(t_out))


def _adaptive_odeint(f, k1, x, dt, t_span, solver, atol=1e-4, rtol=1e-4, args=None, interpolator=None,
                     return_all_eval=False, seminorm=(False, None)):
    """Adaptive loop (reference odeint.py:326-408) with the scalar bookkeeping on the host in fp32."""
    ts = _host_times(t_span)
    t_eval, t, T = ts[1:], f32(ts[0]), f32(ts[-1])
    dt = f32(dt.item() if isinstance(dt, Tensor) else dt)
    if type(solver) in BUILTIN_TYPES:
        attempt = _KernelAttempt(solver, x, atol, rtol, seminorm, interpolator)
    else:
        attempt = _TorchAttempt(solver, x, atol, rtol, seminorm, interpolator, args)
    safety, min_factor, max_factor = solver.controller_constants()
    ck, ck_flag, dt_keep = 0, False, None
    times, sol = [t], [x]
    n_attempts = 0
    trace = None
    if TRACE_SINK is not None:
        trace = {"dt0": float(dt), "t": [], "dt": [], "ratio": [], "accept": []}
        TRACE_SINK.append(trace)
    while t < T:
        n_attempts += 1
        if n_attempts > MAX_STEPS or np.isnan(dt):
            raise RuntimeError(f"torchdyn_b200.odeint: adaptive loop did not terminate (t={t}, dt={dt}, "
                               f"{n_attempts - 1} attempts); the error estimate is probably NaN")
        if f32(t + dt) > T:
            dt = f32(T - t)
        if ck < len(t_eval) and f32(t + dt) > t_eval[ck] and interpolator is None:
            dt_keep, ck_flag = dt, True          # clamp to the next output time, remember the intended dt
            dt = f32(t_eval[ck] - t)

        f_new, x_new, ks, ratio = attempt(f, x, t, dt, k1)
        if trace is not None:
            trace["t"].append(float(t)); trace["dt"].append(float(dt))
            trace["ratio"].append(float(ratio)); trace["accept"].append(bool(ratio <= 1))

        if ratio <= 1:
            t_next = f32(t + dt)
            if interpolator is not None:
                while ck < len(t_eval) and t_next > t_eval[ck]:
                    sol.append(attempt.dense(x, x_new, ks, t, dt, f32(t_eval[ck])))
                    times.append(f32(t_eval[ck]))
                    ck += 1
            hit = ck < len(t_eval) and t_next == t_eval[ck]     # exact float equality, like the reference
            if hit or return_all_eval:
                sol.append(x_new)
                times.append(t_next)
                if hit:
                    ck += 1
            t, x, k1 = t_next, x_new, f_new

        if ck_flag:
            dt = f32(dt_keep - dt)               # the REMAINDER of the intended step (odeint.py:399-401)
            ck_flag = False
        dt = adapt_step(dt, ratio, safety, min_factor, max_factor, solver.order)
    return _times_tensor(times, x), torch.stack(sol)


def _init_step_torch(f, f0, x0, t0, order, atol, rtol):
    """init_step for user solver modules (torch ops, reference arithmetic)."""
    from .utils import hairer_norm
    scale = atol + torch.abs(x0) * rtol
    d0, d1 = hairer_norm(x0 / scale), hairer_norm(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=t0.dtype, device=t0.device)
    else:
        h0 = 0.01 * d0 / d1
    f_new = f(t0 + h0, x0 + h0 * f0)
    d2 = hairer_norm((f_new - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=t0.dtype, device=t0.device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1. / float(order + 1))
    return f32(torch.min(100 * h0, h1).item())


# ----------------------------------------------------------------------------------------------
# fixed-step loop
# ----------------------------------------------------------------------------------------------

def _isclose_count(t, save_at: np.ndarray) -> int:
    # torch.isclose defaults (rtol=1e-5, atol=1e-8), evaluated like ATen: |a-b| <= atol + rtol*|b|
    return int(np.sum(np.abs(f32(t) - save_at) <= (1e-8 + 1e-5 * np.abs(save_at))))


def _fixed_odeint(f, x, t_span, solver, save_at=(), args={}):
    """Fixed-step loop (reference odeint.py:411-445): no FSAL carry, ``dt`` re-derived from the ACCUMULATED t,
    outputs where ``isclose(t, save_at)``.  Built-in solvers run on the CUDA kernels (or, for MLP vector
    fields, on the fused kernels); user solver modules run their own ``step``."""
    ts = _host_times(t_span)
    save_ret = save_at
    if len(save_at) == 0:
        save_ret = None
        sv = ts
    else:
        sv = _host_times(save_at if isinstance(save_at, Tensor) else torch.as_tensor(np.asarray(save_at, dtype=np.float32)))
    assert all(_isclose_count(s, sv) == 1 for s in sv), \
        "each element of save_at [torch.Tensor] must be contained in t_span [torch.Tensor] once and only once"

    builtin = type(solver) in BUILTIN_TYPES and isinstance(x, Tensor)
    proto = x if isinstance(x, Tensor) else x[next(iter(x))]
    if builtin:
        from .. import fused
        out = None if torch.is_grad_enabled() else fused.try_fixed(f, x, ts, sv, solver)
        if out is not None:
            sol = out
        else:
            sol = _fixed_loop_kernels(f, x if x.is_contiguous() else x.contiguous(), ts, sv, solver)
    else:
        sol = _fixed_loop_torch(f, x, ts, sv, solver, args, proto)

    assert len(sol) > 0, \
        "each element of save_at [torch.Tensor] must be contained in t_span [torch.Tensor] once and only once"
    if isinstance(sol[0], dict):
        final_out = {k: torch.stack([s[k] for s in sol]) for k in sol[0].keys()}
    elif isinstance(sol[0], Tensor):
        final_out = torch.stack(sol)
    else:
        raise NotImplementedError(f"{type(x)} is not supported as the state variable")
    if save_ret is None:
        save_ret = _times_tensor(list(ts), proto)
    elif not isinstance(save_ret, Tensor):
        save_ret = torch.tensor(save_ret)
    return save_ret, final_out


def _fixed_loop_kernels(f, x, ts, sv, solver):
    kern = _kernels.get(x)
    p = solver.plan()
    mode_last = 1 if len(p.bsol) > 1 else 0
    t, dt = f32(ts[0]), f32(ts[1] - ts[0])
    sol = [x] if _isclose_count(t, sv) else []
    for step in range(1, len(ts)):
        ks = solver.stages(kern, f, x, t, dt, None)
        x = rk_combine(kern, x, ks[:len(p.bsol)], p.bsol, mode_last, dt)
        t = f32(t + dt)
        if _isclose_count(t, sv):
            sol.append(x)
        if step < len(ts) - 1:
            dt = f32(ts[step + 1] - t)
    return sol


def _fixed_loop_torch(f, x, ts, sv, solver, args, proto):
    mk = lambda v: torch.full((), float(v), dtype=proto.dtype, device=proto.device)  # noqa: E731
    t, dt = f32(ts[0]), f32(ts[1] - ts[0])
    sol = [x] if _isclose_count(t, sv) else []
    for step in range(1, len(ts)):
        _, x, _ = solver.step(f, x, mk(t), mk(dt), k1=None, args=args)
        t = f32(t + dt)
        if _isclose_count(t, sv):
            sol.append(x)
        if step < len(ts) - 1:
            dt = f32(ts[step + 1] - t)
    return sol
