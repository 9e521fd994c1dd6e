"""Natural cubic spline of the forward trajectory for the interpolated adjoint.

The reference delegates this to the third-party ``torchcde`` (``natural_cubic_coeffs`` +
``CubicSpline.evaluate``, call sites torchdyn/numerics/sensitivity.py:126-127, :132; constraint
``torchcde ^0.2.3``, not vendored).  This is a restatement of that published algorithm on two CUDA kernels
(``tdyn_spline_build`` / ``tdyn_spline_eval``): the knot-time-only quantities (1/h, the Thomas multipliers and
pivots) are computed once on the host in fp32, the per-channel sweeps run one thread per channel directly on
the ``[M, B, D]`` trajectory the forward pass saved (the reference's ``permute(1, 0, 2)`` copy is not needed).
Extrapolation clamps the piece index, like ``CubicSpline.evaluate``.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor

from .. import _kernels

f32 = np.float32


class NaturalCubicSpline:
    def __init__(self, sol: Tensor, t_knots):
        """sol: [M, ...] states at the M knot times ``t_knots`` (host array-like or tensor)."""
        self.kern = _kernels.get(sol)
        self.sol = sol if sol.is_contiguous() else sol.contiguous()
        t = t_knots.detach().to("cpu", torch.float32).numpy() if isinstance(t_knots, Tensor) else np.asarray(t_knots, f32)
        self.t = t.astype(f32)
        m = self.sol.shape[0]
        if m < 2:
            raise ValueError("need at least two knots")
        self.kd = torch.empty_like(self.sol)
        self.two_c = torch.empty_like(self.sol[:-1])
        self.three_d = torch.empty_like(self.sol[:-1])
        if m == 2:
            self.kern.spline_build(self.sol, None, self.kd, self.two_c, self.three_d, f32(self.t[1] - self.t[0]))
            return
        h = (self.t[1:] - self.t[:-1]).astype(f32)
        r = (f32(1.0) / h).astype(f32)
        r2 = (r * r).astype(f32)
        diag = np.zeros(m, f32)
        diag[:-1] = r
        diag[1:] += r
        diag *= f32(2.0)
        w = np.zeros(m, f32)
        dp = np.zeros(m, f32)
        dp[0] = diag[0]
        for i in range(1, m):                      # Thomas forward elimination of the (channel-independent) matrix
            w[i] = f32(r[i - 1] / dp[i - 1])
            dp[i] = f32(diag[i] - f32(w[i] * r[i - 1]))
        aux = np.zeros((4, m), f32)
        aux[0, :-1], aux[1, :-1], aux[2], aux[3] = r, r2, w, dp
        self.aux = torch.from_numpy(aux).to(self.sol.device)
        self.kern.spline_build(self.sol, self.aux, self.kd, self.two_c, self.three_d, 0.0)

    def piece(self, t) -> int:
        """clamp(bucketize(t, knots) - 1, 0, M-2): number of knots strictly below t, minus one."""
        idx = int(np.searchsorted(self.t, f32(t), side="left")) - 1
        return min(max(idx, 0), len(self.t) - 2)

    def evaluate(self, t) -> Tensor:
        t = f32(t.item() if isinstance(t, Tensor) else t)
        i = self.piece(t)
        out = torch.empty_like(self.sol[0])
        return self.kern.spline_eval(out, self.sol[i], self.kd[i], self.two_c[i], self.three_d[i], f32(t - self.t[i]))
