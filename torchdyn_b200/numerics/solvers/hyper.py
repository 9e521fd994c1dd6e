"""Hypersolvers: a base explicit step plus a learned correction ``dt**(p+1) * g(t, x)`` (reference
torchdyn/numerics/solvers/hyper.py:17-47).  They are ordinary solver plug-ins: the base step runs on the CUDA kernels
(``super().step``), the correction network is a torch call; ``odeint`` drives them through the plug-in contract."""
import torch

from .ode import Euler, Midpoint, RungeKutta4


class _HyperMixin:
    def _init_hyper(self, hypernet):
        self.hypernet = hypernet
        self.stepping_class = 'fixed'
        self.op1 = self.order + 1

    def step(self, f, x, t, dt, k1=None, args=None):
        _, x_sol, _ = super().step(f, x, t, dt, k1)
        return None, x_sol + dt ** (self.op1) * self.hypernet(t, x), None


class HyperEuler(_HyperMixin, Euler):
    def __init__(self, hypernet, dtype=torch.float32):
        Euler.__init__(self, dtype)
        self._init_hyper(hypernet)


class HyperMidpoint(_HyperMixin, Midpoint):
    def __init__(self, hypernet, dtype=torch.float32):
        Midpoint.__init__(self, dtype)
        self._init_hyper(hypernet)


class HyperRungeKutta4(_HyperMixin, RungeKutta4):
    def __init__(self, hypernet, dtype=torch.float32):
        RungeKutta4.__init__(self, dtype)
        self._init_hyper(hypernet)
