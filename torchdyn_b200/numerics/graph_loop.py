"""Device-side adaptive loop for an OPAQUE vector field: the whole accept / reject loop of ``_adaptive_odeint``
(reference numerics/odeint.py:351-407) as ONE CUDA-graph launch.

The body of the loop -- one attempted step: controller begin, six (stage combine, ``f``) pairs, finish, controller end,
commit -- is captured once with ``torch.cuda.graph`` (``f`` stays a torch call: its kernels, cuDNN included, are captured
by PyTorch; the stage / finish kernels read ``dt`` from the controller block on the device), wrapped by
``tdyn_graph_while_create`` into a conditional WHILE node whose condition the controller kernel sets, and launched once
per ``odeint`` call.  The host then reads the counters with a single synchronisation; the host-driven loop needs one per
attempted step.  Step sizes, accept / reject decisions and outputs follow the same fp32 arithmetic as the host loop
(``numerics/odeint.py::_adaptive_odeint``): traces are identical.

Opt-in (``graph_loop.ENABLED = True`` or ``TDYN_DEVICE_LOOP=1``): a vector field must be capturable (no host
synchronisation, no data-dependent Python control flow); when capture fails the host loop runs instead.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import numpy as np
import torch

from .. import _kernels
from .. import distributed as _dist
from .solvers.ode import BUILTIN_TYPES, Reversed

f32 = np.float32
ENABLED = os.environ.get("TDYN_DEVICE_LOOP", "0") not in ("", "0")
CACHE = False             # re-use a captured loop for the same (f, shapes, tolerances): only safe when f reads no tensor that
                          # is RE-ALLOCATED between calls (parameters updated in place are fine)
MAX_EXTRA_OUT = 2048      # return_all_eval: accepted steps beyond the t_span points that fit the output buffer
TRACE_CAP = 3 * 4096 + 1
MAX_STEPS = 1_000_000
last_status = None        # why the most recent call did (not) take the device loop: "ok" | reason

# controller-block indices (csrc/tdyn_graph_loop.cu)
F_T, F_DT, F_TEND, F_DT_KEEP, F_DT_CUR, F_RATIO, F_SAFETY, F_MINF, F_MAXF, F_C0, F_TS0 = 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 16
(I_CK, I_CK_FLAG, I_N_EVAL, I_N_ATTEMPT, I_N_ACCEPT, I_N_OUT, I_OUT_CAP, I_RETURN_ALL, I_KEEP_GOING, I_ACCEPT, I_COPY_SLOT,
 I_STATUS, I_MAX_STEPS, I_ORDER, I_N_STAGE_T, I_TRACE_CAP) = range(16)

_cache = weakref.WeakKeyDictionary()


def _tensor_callable(f_):
    """``f_`` as a callable of a DEVICE time tensor, or None when it can only be driven with host times: the library's own
    dynamics objects (fused fields, adjoint dynamics) take host-side times -- and an adjoint's VJP is issued by the autograd
    engine's threads from inside a running backward pass, which stream capture does not survive (measured) -- so they keep
    the host-driven loop."""
    inner = f_.f if isinstance(f_, Reversed) else f_
    if hasattr(inner, "tdyn_eval"):
        return None
    return f_


def _counters(f_):
    """Objects along the wrapper chain whose ``nfe`` advances per evaluation (DEFunc / DEFuncBase / MLPField)."""
    seen, out, stack = set(), [], [f_]
    while stack and len(seen) < 16:
        o = stack.pop()
        if id(o) in seen or o is None:
            continue
        seen.add(id(o))
        if isinstance(getattr(o, "nfe", None), (int, float)):
            out.append(o)
        for name in ("f", "vf"):
            nxt = getattr(o, name, None)
            if nxt is not None and not isinstance(nxt, torch.Tensor):
                stack.append(nxt)
    return out


def _recover_after_failed_capture(device):
    """A capture that dies half way leaves PyTorch's CUDA generator flagged as "capturing" (its epilogue never ran), which
    would make the next random-number call fail: an empty capture runs prologue + epilogue and clears the flag."""
    import warnings
    try:
        torch.cuda.synchronize(device)
    except Exception:
        pass
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                pass
    except Exception:
        pass


class _Loop:
    """One captured + instantiated WHILE graph and its static buffers."""

    def __init__(self, kern, fn, x, plan, solver, atol, rtol, n_eval, return_all, norm_n, norm_count, counters):
        dev = x.device
        self.kern, self.n_eval = kern, n_eval
        lib = kern.lib
        import ctypes as C
        nf, ni = C.c_int(), C.c_int()
        lib.tdyn_loop_ctrl_sizes(C.byref(nf), C.byref(ni))
        self.cf = torch.zeros(nf.value, dtype=torch.float32, device=dev)
        self.ci = torch.zeros(ni.value, dtype=torch.int32, device=dev)
        self.cf_host = torch.zeros(nf.value, dtype=torch.float32).pin_memory()
        self.ci_host = torch.zeros(ni.value, dtype=torch.int32).pin_memory()
        self.cap = 1 + n_eval + (MAX_EXTRA_OUT if return_all else 0)
        self.t_eval = torch.zeros(max(n_eval, 1), dtype=torch.float32, device=dev)
        self.t_out = torch.zeros(self.cap, dtype=torch.float32, device=dev)
        self.trace = torch.zeros(TRACE_CAP, dtype=torch.float32, device=dev)
        self.xs = torch.empty_like(x)
        self.x_new = torch.empty_like(x)
        self.k1 = torch.empty_like(x)
        ys = [torch.empty_like(x) for _ in plan.stages]
        tviews = [self.cf[F_TS0 + i:F_TS0 + i + 1] for i in range(len(plan.stages))]
        dt_ptr = self.cf.data_ptr() + 4 * F_DT_CUR
        before = [c.nfe for c in counters]
        self.graph = torch.cuda.CUDAGraph(keep_graph=True)
        # accepted states are committed into this loop-owned buffer; the caller gets a copy of the rows that were written
        self.out = torch.empty(self.cap, *x.shape, dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        # (capture mode "global": the VJP of an adjoint's dynamics is launched by the autograd engine's device thread)
        with torch.cuda.graph(self.graph):
            kern.loop_begin(self.cf, self.ci, self.t_eval)
            ks = [self.k1]
            for i, st in enumerate(plan.stages):
                kern.combine_dev(ys[i], self.xs, ks[:len(st.coefs)], st.coefs, st.mode, dt_ptr)
                k = fn(tviews[i], ys[i])
                if k.dtype != x.dtype:
                    k = k.to(x.dtype)
                ks.append(k if k.is_contiguous() else k.contiguous())
            kern.finish_dev(self.x_new, self.xs, ks, plan.bsol, plan.berr, dt_ptr, atol, rtol, norm_n)
            kern.loop_end(self.cf, self.ci, norm_count, self.t_eval, self.t_out, self.trace)
            kern.loop_commit(self.ci, self.xs, self.x_new, self.k1, ks[-1], self.out)
        self._keep = (ys, ks)
        self.nfe_per_attempt = [c.nfe - b for c, b in zip(counters, before)]
        for c, b in zip(counters, before):      # the capture itself evaluated nothing
            c.nfe = b
        self.exec = kern.graph_while_create(self.graph.raw_cuda_graph(), self.ci.data_ptr() + 4 * I_KEEP_GOING)

    def __del__(self):
        try:
            if getattr(self, "exec", None):
                self.kern.graph_destroy(self.exec)
        except Exception:
            pass

    def run(self, x, k1, t0, T, dt0, t_eval, solver, return_all):
        kern = self.kern
        safety, min_factor, max_factor = solver.controller_constants()
        c = [float(st.c) for st in solver.plan().stages]
        cf, ci = self.cf_host, self.ci_host
        cf.zero_(); ci.zero_()
        cf[F_T], cf[F_DT], cf[F_TEND] = float(t0), float(dt0), float(T)
        cf[F_SAFETY], cf[F_MINF], cf[F_MAXF] = float(safety), float(min_factor), float(max_factor)
        for i, v in enumerate(c):
            cf[F_C0 + i] = v
        ci[I_N_EVAL], ci[I_N_OUT], ci[I_OUT_CAP], ci[I_RETURN_ALL] = self.n_eval, 1, self.cap, int(bool(return_all))
        ci[I_KEEP_GOING], ci[I_MAX_STEPS], ci[I_ORDER], ci[I_N_STAGE_T], ci[I_TRACE_CAP] = 1, MAX_STEPS, solver.order, len(c), TRACE_CAP
        self.cf.copy_(cf, non_blocking=True)
        self.ci.copy_(ci, non_blocking=True)
        if self.n_eval:
            self.t_eval.copy_(torch.from_numpy(np.ascontiguousarray(t_eval, dtype=np.float32)), non_blocking=True)
        self.xs.copy_(x)
        self.k1.copy_(k1)
        self.out[0].copy_(x)
        self.t_out[:1].fill_(float(t0))
        self.trace[:1].fill_(float(dt0))
        kern.graph_launch(self.exec)
        self.ci_host.copy_(self.ci, non_blocking=True)
        torch.cuda.current_stream(x.device).synchronize()          # the ONE host synchronisation of the solve
        return self.ci_host.tolist()


def try_device_loop(f_user, f_, k1, x, dt0, ts, solver, atol, rtol, interpolator, return_all_eval, seminorm) -> Optional[tuple]:
    """Run the adaptive loop on the device; ``None`` -> the caller runs the host-driven loop."""
    global last_status
    if not ENABLED:
        last_status = "disabled"
        return None
    if type(solver) not in BUILTIN_TYPES or solver.stepping_class != "adaptive" or interpolator is not None:
        last_status = "solver / interpolator not supported"
        return None
    if torch.is_grad_enabled() or _dist.is_enabled() or not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        last_status = "needs no_grad, one GPU, contiguous fp32 CUDA state"
        return None
    fn = _tensor_callable(f_)
    if fn is None:
        last_status = "library-internal dynamics with host-side times"
        return None
    from .odeint import _seminorm_count
    kern = _kernels.get(x)
    if not getattr(kern, "is_cuda", False):
        return None
    t0, T = f32(ts[0]), f32(ts[-1])
    if not (t0 < T) or np.isnan(dt0):
        return None
    t_eval = np.asarray(ts[1:], dtype=np.float32)
    plan = solver.plan()
    norm_n = _seminorm_count(x, seminorm)
    counters = _counters(f_)
    params = tuple(p.data_ptr() for p in f_user.parameters()) if isinstance(f_user, torch.nn.Module) else ()
    key = (tuple(x.shape), x.device.index, type(solver), float(atol), float(rtol), len(t_eval), bool(return_all_eval), norm_n,
           isinstance(f_, Reversed), params)
    loop = None
    holder = None
    if CACHE:
        try:
            holder = _cache.setdefault(f_user, {})
            loop = holder.get(key)
        except TypeError:
            holder = None
    if loop is None:
        try:
            loop = _Loop(kern, fn, x, plan, solver, atol, rtol, len(t_eval), return_all_eval, norm_n, norm_n, counters)
        except Exception as e:      # f is not capturable (host sync, unsupported op ...): host loop
            last_status = f"capture failed: {type(e).__name__}: {str(e).splitlines()[0]}"
            _recover_after_failed_capture(x.device)
            return None
        if holder is not None:
            if len(holder) > 4:
                holder.clear()
            holder[key] = loop
    st = loop.run(x, k1, t0, T, dt0, t_eval, solver, return_all_eval)
    n_attempt, n_out, status = st[I_N_ATTEMPT], st[I_N_OUT], st[I_STATUS]
    if status == 2:
        last_status = "output buffer full"
        return None
    if status == 1:
        raise RuntimeError(f"torchdyn_b200.odeint: adaptive loop did not terminate ({n_attempt} attempts); the error "
                           "estimate is probably NaN")
    for c, d in zip(counters, loop.nfe_per_attempt):      # the captured body evaluated f on the device n_attempt times
        c.nfe += d * n_attempt
    from . import odeint as _  # noqa: F401
    import sys
    om = sys.modules["torchdyn_b200.numerics.odeint"]
    if om.TRACE_SINK is not None:
        tr = loop.trace[:3 * min(n_attempt, (TRACE_CAP - 1) // 3) + 1].cpu().tolist()
        rec = {"dt0": tr[0], "t": tr[1::3], "dt": tr[2::3], "ratio": tr[3::3]}
        rec["accept"] = [r <= 1 for r in rec["ratio"]]
        om.TRACE_SINK.append(rec)
    last_status = "ok"
    return loop.t_out[:n_out].clone(), loop.out[:n_out].clone()
