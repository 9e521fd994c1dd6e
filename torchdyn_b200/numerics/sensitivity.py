"""Adjoint sensitivities as ``torch.autograd.Function`` factories.

Reference: torchdyn/numerics/sensitivity.py:32-100 (backsolve adjoint: augmented flat state
``[x, lam, mu, lam_t]`` integrated in reverse per ``t_span`` interval with the seminorm over ``[x, lam]``)
and :104-161 (interpolated adjoint: ``[lam, mu]`` with x(t) from a natural cubic spline of ALL accepted
forward steps; forward solver / tolerances, no seminorm).

Both factories keep the reference's autograd contract -- ``forward(ctx, vf_params, x, t_span, B=None,
save_at=())`` -> ``(t_sol, sol)``; ``backward`` -> ``(mu, lam, lam_tspan | None, None, ...)`` -- so gradients
reach the parameters through the flat ``vf_params`` input.  The integration itself (forward and reverse) runs
on the CUDA kernels of ``numerics.odeint``; the augmented dynamics evaluates the vector field's VJP with
torch autograd for arbitrary modules.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.autograd import Function, grad

from .. import distributed as _dist
from .odeint import odeint
from .spline import NaturalCubicSpline

f32 = np.float32


def generic_odeint(problem_type, vf, x, t_span, solver, atol, rtol, interpolator, B0=None,
                   return_all_eval=False, maxiter=4, fine_steps=4, save_at=()):
    "Dispatch on the problem class; only 'standard' (ODEProblem) is part of this implementation."
    if problem_type == 'standard':
        return odeint(vf, x, t_span, solver, atol=atol, rtol=rtol, interpolator=interpolator,
                      return_all_eval=return_all_eval, save_at=save_at)
    raise NotImplementedError(f"problem_type {problem_type!r} (multiple shooting) is outside torchdyn_b200's scope")


def _time(t_host, like):
    return torch.full((), float(t_host), dtype=like.dtype, device=like.device)


def _vjp(vf, t, x, lam, integral_loss):
    """(f, -lam^T df/dx [- dg/dx], -lam^T df/dtheta flattened, -lam^T df/dt)  (sensitivity.py:62-74)."""
    with torch.set_grad_enabled(True):
        x, t = x.detach().requires_grad_(True), t.detach().requires_grad_(True)
        dx = vf(t, x)
        dlam, dt_, *dmu = tuple(grad(dx, (x, t) + tuple(vf.parameters()), -lam, allow_unused=True,
                                     retain_graph=False))
        if integral_loss:
            dg = torch.autograd.grad(integral_loss(t, x).sum(), x, allow_unused=True, retain_graph=True)[0]
            dlam = dlam - dg
        dmu = torch.cat([e.flatten() if e is not None else torch.zeros(1, dtype=x.dtype, device=x.device) for e in dmu],
                        dim=-1)
        if dt_ is None:
            dt_ = torch.zeros(1, dtype=t.dtype, device=t.device)
        if dt_.dim() == 0:
            dt_ = dt_.unsqueeze(0)
    return dx.detach(), dlam, dmu, dt_


def _integral_grad(integral_loss, t, x):
    """dg/dx of the running cost g(t, x) (reference sensitivity.py:68, :142)."""
    with torch.set_grad_enabled(True):
        x = x.detach().requires_grad_(True)
        dg = torch.autograd.grad(integral_loss(t, x).sum(), x, allow_unused=True)[0]
    return torch.zeros_like(x) if dg is None else dg


class BacksolveDynamics:
    """d/dt [x, lam, mu, lam_t] for the backsolve adjoint on the flat vector A (sensitivity.py:57-75)."""

    def __init__(self, vf, x_shape, lam_shape, n_params, integral_loss, device=None):
        self.vf, self.x_shape, self.lam_shape, self.P, self.integral_loss = vf, x_shape, lam_shape, n_params, integral_loss
        self.n = int(np.prod(x_shape))
        self.fused = None
        if device is not None and tuple(x_shape) == tuple(lam_shape):
            from .. import fused
            self.fused = fused.adjoint_for(vf, tuple(x_shape), device)

    def tdyn_eval(self, t_host, A, sign: float = 1.0):
        if self.fused is None or not A.is_contiguous():
            out = self(_time(t_host, A), A)
            return out if sign == 1.0 else -out
        # MLP field: f, -lam^T df/dx and the batch-reduced -lam^T df/dtheta in ONE kernel, written straight into the
        # flat [dx, dlam, dmu, dlam_t] layout (df/dt = 0 for an autonomous MLP -> dlam_t = 0, sensitivity.py:73)
        n, P = self.n, self.P
        out = torch.empty_like(A)
        out[-1:].zero_()
        self.fused(A[:n], A[n:2 * n], out[:n], out[n:2 * n], out[2 * n:2 * n + P], self.x_shape[0], sign)
        if self.integral_loss:
            # integral loss: the co-state dynamics gain -dg/dx (sensitivity.py:67-69); g is a user callable, so its
            # gradient is a torch call, added to the fused kernel's output
            out[n:2 * n].sub_(_integral_grad(self.integral_loss, _time(t_host, A), A[:n].reshape(self.x_shape)).flatten(), alpha=sign)
        return out

    def __call__(self, t, A):
        if t.dim() > 0:
            t = t[0]
        n = self.n
        x, lam = A[:n].reshape(self.x_shape), A[n:2 * n].reshape(self.lam_shape)
        dx, dlam, dmu, dt_ = _vjp(self.vf, t, x, lam, self.integral_loss)
        return torch.cat([dx.flatten(), dlam.flatten(), dmu.flatten(), dt_.flatten()])


class InterpDynamics:
    """d/dt [lam, mu] with x(t) from the spline (sensitivity.py:130-147)."""

    def __init__(self, vf, spline, lam_shape, n_params, integral_loss, device=None):
        self.vf, self.spline, self.lam_shape, self.P, self.integral_loss = vf, spline, lam_shape, n_params, integral_loss
        self.n = int(np.prod(lam_shape))
        self.fused = None
        if device is not None:
            from .. import fused
            self.fused = fused.adjoint_for(vf, tuple(lam_shape), device)

    def tdyn_eval(self, t_host, A, sign: float = 1.0):
        x = self.spline.evaluate(t_host)
        if self.fused is not None and A.is_contiguous() and x.is_contiguous():
            n, P = self.n, self.P
            out = torch.empty_like(A)
            # f(x) itself is not part of [dlam, dmu]: kernels that can skip it take None
            scratch = None if getattr(self.fused, "f_optional", False) else torch.empty_like(x)
            self.fused(x, A[:n], scratch, out[:n], out[n:n + P], self.lam_shape[0], sign)
            if self.integral_loss:      # -dg/dx on the co-state (sensitivity.py:141-143)
                out[:n].sub_(_integral_grad(self.integral_loss, _time(t_host, A), x.reshape(self.lam_shape)).flatten(), alpha=sign)
        else:
            out = self._eval_torch(t_host, A, x)
            out = out if sign == 1.0 else -out
        if _dist.is_enabled():
            # batch sharded over ranks: d(mu)/dt is a sum over the batch -> add the per-rank partials so that mu (which
            # sits INSIDE the error norm here, sensitivity.py:152) is the same global vector on every rank
            _dist.allreduce_sum_(out[self.n:self.n + self.P])
        return out

    def _eval_torch(self, t_host, A, x):
        lam = A[:self.n].reshape(self.lam_shape)
        _, dlam, dmu, _ = _vjp(self.vf, _time(t_host, A), x, lam, self.integral_loss)
        return torch.cat([dlam.flatten(), dmu.flatten()])

    def __call__(self, t, A):
        return self.tdyn_eval(f32(t.reshape(-1)[0].item()), A)


def _gather_odefunc_adjoint(vf, vf_params, solver, atol, rtol, interpolator, solver_adjoint,
                            atol_adjoint, rtol_adjoint, integral_loss, problem_type, maxiter=4, fine_steps=4):
    "autograd.Function for the backsolve adjoint of an ODEProblem"

    class _ODEProblemFunc(Function):
        @staticmethod
        def forward(ctx, vf_params, x, t_span, B=None, save_at=()):
            t_sol, sol = generic_odeint(problem_type, vf, x, t_span, solver, atol, rtol, interpolator, B,
                                        False, maxiter, fine_steps, save_at)
            ctx.save_for_backward(sol, t_sol)
            ctx.t_host = t_sol.detach().to("cpu", torch.float32).numpy()
            return t_sol, sol

        @staticmethod
        def backward(ctx, *grad_output):
            sol, t_sol = ctx.saved_tensors
            th = ctx.t_host
            g_t, g_sol = grad_output[0], grad_output[-1]
            P = sum(p.numel() for p in vf.parameters())
            xT, lamT = sol[-1], g_sol[-1]
            n = xT.numel()
            from .. import fused
            lam_flat = lamT.flatten()
            lam_t = lam_flat @ fused.eval_field(vf, t_sol[-1], xT).flatten()                 # :53
            A = torch.cat([xT.flatten(), lam_flat, torch.zeros(P, dtype=xT.dtype, device=xT.device), lam_t[None]])
            dyn = BacksolveDynamics(vf, xT.shape, lamT.shape, P, integral_loss, device=xT.device)

            # (the reference fills a dLdt vector over all of t_sol and hands back its two end points, sensitivity.py:78-98: only
            # those two are kept here, and the last interval does not re-assemble a flat A that nobody integrates)
            dLdt_last, dLdt_first = lam_t, lam_t
            lam = mu = None
            for i in range(len(th) - 1, 0, -1):
                last = fused.try_adjoint_interval(dyn, A, th[i], th[i - 1], solver_adjoint, atol_adjoint, rtol_adjoint)
                if last is None:
                    _, As = odeint(dyn, A, np.array([th[i], th[i - 1]], dtype=np.float32), solver_adjoint,
                                   atol=atol_adjoint, rtol=rtol_adjoint, seminorm=(True, 2 * n))
                    last = As[-1]
                xt = last[:n].reshape(xT.shape)
                # (evaluated at every interval like the reference does -- it is an NFE there too; index i: reference quirk, :88)
                dLdt_first = last[n:2 * n] @ fused.eval_field(vf, t_sol[i], xt.contiguous()).flatten()
                if i == 1:
                    lam = (last[n:2 * n] + g_sol[0].flatten()).reshape(lamT.shape)
                    mu = last[-P - 1:-1]
                else:
                    lt = last[-1:] - g_t[i - 1]
                    A = torch.cat([last[:n], last[n:2 * n] + g_sol[i - 1].flatten(), last[-P - 1:-1], lt])
            return (mu, lam, torch.stack([dLdt_first, dLdt_last]), None, None, None)

    return _ODEProblemFunc


def _gather_odefunc_interp_adjoint(vf, vf_params, solver, atol, rtol, interpolator, solver_adjoint,
                                   atol_adjoint, rtol_adjoint, integral_loss, problem_type, maxiter=4, fine_steps=4):
    "autograd.Function for the interpolated adjoint of an ODEProblem"

    class _ODEProblemFunc(Function):
        @staticmethod
        def forward(ctx, vf_params, x, t_span, B=None, save_at=()):
            t_sol, sol = generic_odeint(problem_type, vf, x, t_span, solver, atol, rtol, interpolator, B,
                                        True, maxiter, fine_steps, save_at)
            ctx.save_for_backward(sol, t_sol)
            ctx.t_host = t_sol.detach().to("cpu", torch.float32).numpy()
            ctx.span_host = (t_span.detach().to("cpu", torch.float32).numpy() if isinstance(t_span, torch.Tensor)
                             else np.asarray(t_span, dtype=np.float32))
            return t_sol, sol

        @staticmethod
        def backward(ctx, *grad_output):
            sol, t_sol = ctx.saved_tensors
            span = ctx.span_host
            g_sol = grad_output[-1]
            P = sum(p.numel() for p in vf.parameters())
            lamT = g_sol[-1]
            n = lamT.numel()
            A = torch.cat([lamT.flatten(), torch.zeros(P, dtype=lamT.dtype, device=lamT.device)])
            spline = NaturalCubicSpline(sol.detach(), ctx.t_host)
            dyn = InterpDynamics(vf, spline, lamT.shape, P, integral_loss, device=lamT.device)
            from .. import fused
            solver_obj = solver if not isinstance(solver, str) else None
            with _dist.replicated_tail(n):          # sharded batch: lam is per-rank, mu is replicated (counted once)
                for i in range(len(span) - 1, 0, -1):
                    # forward solver and tolerances, no seminorm (sensitivity.py:152)
                    last = None
                    if solver_obj is not None:
                        last = fused.try_interp_interval(dyn, A, span[i], span[i - 1], solver_obj, atol, rtol)
                    if last is None:
                        _, As = odeint(dyn, A, np.array([span[i], span[i - 1]], dtype=np.float32), solver, atol=atol, rtol=rtol)
                        last = As[-1]
                    A = torch.cat([last[:n] + g_sol[i - 1].flatten(), last[-P:]])
            lam, mu = A[:n].reshape(lamT.shape), A[-P:]
            if _dist.is_enabled() and _dist.rank() != 0:
                mu = torch.zeros_like(mu)           # mu is already the global sum: hand it out once, so that the
                                                    # caller's allreduce_gradients (sum over ranks) stays correct
            return (mu, lam, None, None, None)

    return _ODEProblemFunc
