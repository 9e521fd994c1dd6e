"""Norm, initial step and step-size controller of the adaptive loop.

Reference: torchdyn/numerics/utils.py:33-34 (hairer_norm), :37-54 (init_step), :57-63 (adapt_step).
The reductions run in CUDA (``tdyn_init_norms``); the scalar controller arithmetic runs on the host in
IEEE fp32 (numpy float32 scalars) -- the same roundings the reference gets from 0-dim fp32 tensors, without
its ~7 device->host syncs per step.  When the batch is sharded over GPUs the per-GPU sums of squares are
added across ranks before the square root (see ``torchdyn_b200.distributed``), so every rank takes the
decisions of the reference's single global controller.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _kernels
from .. import distributed as _dist
from .._lib import MODE_CHAIN
from .solvers.ode import call_vf

f32 = np.float32


def hairer_norm(tensor: torch.Tensor) -> torch.Tensor:
    """sqrt(mean(|v|^2)) over all elements -- public helper kept for API compatibility (torch ops; the
    integration loops use the fused CUDA reductions instead)."""
    return tensor.abs().pow(2).mean().sqrt()


def rms_from_sumsq(sumsq: float, count: int, kern=None) -> "np.float32":
    """error_ratio = sqrt(mean) in fp32 from the fp64 sum of fp32 squares.  The square root is the correctly rounded
    IEEE one -- what the reference's ``.sqrt()`` is on a CUDA tensor.  (torch's CPU ``sqrt`` is NOT correctly rounded in
    ~0.6 % of the cases; the CPU test double of the kernel backend supplies ``sqrt32`` = torch's CPU sqrt so that the host
    logic can be compared bit for bit with the reference executed on CPU.)"""
    return getattr(kern, "sqrt32", np.sqrt)(f32(sumsq / float(count)))


def init_step(f, f0, x0, t0, order, atol, rtol):
    """Hairer's starting step size (utils.py:37-54).  Returns an fp32 host scalar.

    NB (reference quirk, kept): ``odeint`` hands over the UN-negated ``f`` together with the negated
    ``t0`` on reversed spans."""
    kern = _kernels.get(x0)
    t0 = f32(t0.item() if isinstance(t0, torch.Tensor) else t0)
    x0, f0 = x0.detach(), f0.detach()       # step sizes are constants of the autograd graph
    cn = _dist.local_norm_count(x0.numel())     # < numel only for a sharded state with a replicated tail (rank != 0)
    nv = (lambda v: v) if cn == x0.numel() else (lambda v: v.reshape(-1)[:cn])
    kern.init_norms(nv(x0), nv(f0), None, atol, rtol)
    (s0, s1), count = _dist.reduce_sums(kern, 2, cn)
    d0, d1 = rms_from_sumsq(s0, count, kern), rms_from_sumsq(s1, count, kern)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = f32(1e-6)
    else:
        h0 = f32(f32(0.01) * d0) / d1
    x_new = kern.combine(kern.empty_like(x0), x0, [f0], (1.0,), MODE_CHAIN, h0)     # x0 + h0*f0
    with torch.no_grad():
        f_new = call_vf(kern, f, f32(t0 + h0), x_new)
    kern.init_norms(nv(x0), nv(f0), nv(f_new), atol, rtol)
    s2 = _dist.reduce_sums(kern, 3, cn)[0][2]
    d2 = f32(rms_from_sumsq(s2, count, kern) / h0)
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(f32(1e-6), f32(h0 * f32(1e-3)))
    else:
        # `tensor ** python_float` on a CPU fp32 tensor is evaluated as double pow(double(base), exponent) and
        # rounded once to fp32 (ATen pow_tensor_scalar); do the same.
        # NB `0.01 / tensor` is Tensor.__rtruediv__ = tensor.reciprocal() * 0.01: two fp32 roundings.
        base = f32(f32(f32(1.0) / max(d1, d2)) * f32(0.01))
        h1 = f32(math.pow(float(base), 1.0 / float(order + 1)))
    return f32(min(f32(f32(100) * h0), h1))


def adapt_step(dt, error_ratio, safety, min_factor, max_factor, order):
    """Step-size controller (utils.py:57-63) on fp32 host scalars: ratio==0 -> dt*max_factor; ratio<1 lifts
    min_factor to 1; factor = min(max_factor, max(safety/ratio**(1/order), min_factor))."""
    dt, r = f32(dt), f32(error_ratio)
    if r == 0:
        return f32(dt * f32(max_factor))
    lo = f32(1.0) if r < 1 else f32(min_factor)
    expo = f32(1.0) / f32(order)
    with np.errstate(all="ignore"):
        cand = f32(f32(safety) / f32(r ** expo))      # numpy scalar ** == libm powf == ATen's 0-dim tensor ** tensor
        # torch.max / torch.min propagate NaN
        factor = cand if np.isnan(cand) else min(f32(max_factor), max(cand, lo))
    return f32(dt * factor)
