"""TEST DOUBLE for the CUDA kernel backend: the same entry points as torchdyn_b200._kernels.CudaKernels,
computed with torch-CPU fp32 ops in the reference's association order (the checker's arithmetic).  It lets the
CPU test-suite drive the PRODUCT's host logic (loops, controller, adjoint plumbing, sharding) without a GPU.
It lives in tests/ and is never importable from the package."""
import numpy as np
import torch

from oracle import erk_oracle as eo


class CpuKernels:
    is_cuda = False

    def __init__(self):
        self.red = torch.zeros(4, dtype=torch.float64)
        self.launches = 0

    @staticmethod
    def _sum_like_torch(v):
        """fp64 value S such that float32(S / numel) == v.abs().pow(2).mean() as torch-CPU computes it in fp32
        (exact while numel < 2**29): lets the host controller reproduce the reference's fp32 mean bit for bit."""
        return v.abs().pow(2).mean().double() * v.numel()

    @staticmethod
    def sqrt32(v):
        """torch's CPU fp32 sqrt (what the executed reference used; differs from the IEEE sqrt in ~0.6 % of the cases)"""
        return np.float32(torch.tensor(float(v), dtype=torch.float32).sqrt().item())

    def empty_like(self, x):
        return torch.empty_like(x)

    def time_tensor(self, value, like):
        return torch.full((1,), float(value), dtype=like.dtype)

    def combine(self, y, x, ks, coefs, mode, dt):
        dt = torch.tensor(float(dt), dtype=torch.float32)
        cf = [torch.tensor(float(c), dtype=torch.float32) for c in coefs]
        if mode == 0:
            out = eo._chain(x, dt, cf, ks)
        else:
            out = eo._inner(x, dt, cf, ks)
        y.copy_(out)
        self.launches += 1
        return y

    def finish(self, x_new, err_out, x, ks, bsol, berr, dt, atol, rtol, norm_n):
        dt = torch.tensor(float(dt), dtype=torch.float32)
        xs = eo._inner(x, dt, [torch.tensor(float(c)) for c in bsol], ks[:len(bsol)])
        err = eo._inner(None, dt, [torch.tensor(float(c)) for c in berr], ks[:len(berr)])
        x_new.copy_(xs)
        if err_out is not None:
            err_out.copy_(err)
        es = (err / (atol + rtol * torch.max(x.abs(), xs.abs()))).flatten()[:norm_n]
        self.red[0] = self._sum_like_torch(es)
        self.launches += 1

    def init_norms(self, x0, f0, f1, atol, rtol):
        scale = atol + torch.abs(x0) * rtol
        self.red[0] = self._sum_like_torch(x0 / scale)
        self.red[1] = self._sum_like_torch(f0 / scale)
        if f1 is not None:
            self.red[2] = self._sum_like_torch((f1 - f0) / scale)
        self.launches += 1

    def dense_eval(self, out, x0, x1, ks, bmid, dt, th):
        dt = torch.tensor(float(dt), dtype=torch.float32)
        bm = [torch.tensor(float(b)) for b in bmid]
        x_mid = x0 + dt * sum([bm[i] * ks[i] for i in range(7)])
        c = eo.dense_fit_4th(dt, ks[0], ks[6], x0, x1, x_mid)
        t = [torch.tensor(float(v), dtype=torch.float32) for v in th]
        r = c[0] + t[0] * c[1]
        r = r + t[1] * c[2]
        r = r + t[2] * c[3]
        r = r + t[3] * c[4]
        out.copy_(r)
        self.launches += 1
        return out

    def spline_build(self, sol, aux, kd, two_c, three_d, dt01):
        raise NotImplementedError("spline double is built in spline_from_oracle")

    def read_reduction(self, n=1):
        return self.red[:n].tolist()


class OracleSpline:
    """stand-in for torchdyn_b200.numerics.spline.NaturalCubicSpline on CPU"""

    def __init__(self, sol, t_knots):
        t = torch.as_tensor(np.asarray(t_knots), dtype=torch.float32)
        flat = sol.reshape(sol.shape[0], -1)            # [M knots, channels]
        co = eo.natural_cubic_coeffs(flat, t)
        self.sp = eo.NaturalCubicSpline.from_coeffs(co, t)
        self.shape = sol.shape[1:]

    def evaluate(self, t):
        return self.sp.evaluate(torch.tensor(float(t))).reshape(self.shape)
