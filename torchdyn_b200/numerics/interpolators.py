"""Dense output for dopri5 (reference: torchdyn/numerics/interpolators.py:29-37, :57-63, registry :67-73).

``FourthOrder`` keeps the reference's object contract (``bmid``, ``fit``, ``evaluate``,
``sync_device_dtype``) so user code can call it, but ``odeint`` evaluates it with ONE fused kernel per
output time (``tdyn_dense_eval``: x_mid, the five polynomial coefficients and the Horner-like evaluation in
registers, 9 reads + 1 write per element instead of ~40 element-wise launches)."""
from __future__ import annotations

import torch

from .solvers._constants import construct_4th


class Interpolator:
    def __init__(self, order):
        self.order = order
        self.bmid = None

    def sync_device_dtype(self, x, t_span):
        if self.bmid is not None:
            self.bmid = self.bmid.to(x)
        return x, t_span

    def fit(self, dt, f0, f1, x0, x1, x_mid, **kwargs):
        raise NotImplementedError

    def evaluate(self, coefs, t0, t1, t):
        "Polynomial in theta = (t - t0)/(t1 - t0) with coefficients ordered from degree 0 upwards."
        theta = (t - t0) / (t1 - t0)
        out = coefs[0] + theta * coefs[1]
        power = theta
        for cf in coefs[2:]:
            power = power * theta
            out = out + power * cf
        return out


class FourthOrder(Interpolator):
    def __init__(self, dtype=torch.float32):
        super().__init__(order=4)
        self.bmid = construct_4th(dtype)

    def fit(self, dt, f0, f1, x0, x1, x_mid, **kwargs):
        "Quartic through (x0, f0), (x1, f1) and the mid-point value; returns [c0..c4] (degree ascending)."
        q4 = 2 * dt * (f1 - f0) - 8 * (x1 + x0) + 16 * x_mid
        q3 = dt * (5 * f0 - 3 * f1) + 18 * x0 + 14 * x1 - 32 * x_mid
        q2 = dt * (f1 - 4 * f0) - 11 * x0 - 5 * x1 + 16 * x_mid
        return [x0, dt * f0, q2, q3, q4]


INTERP_DICT = {'4th': FourthOrder}


def str_to_interp(solver_name, dtype=torch.float32):
    "Interpolation scheme name -> instance."
    return INTERP_DICT[solver_name](dtype)
