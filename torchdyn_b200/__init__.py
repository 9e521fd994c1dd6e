"""torchdyn_b200 -- B200-native drop-in for torchdyn's batched explicit-RK ODE integration path.

    from torchdyn_b200.numerics import odeint           # == torchdyn.numerics.odeint
    from torchdyn_b200.core import NeuralODE, ODEProblem  # == torchdyn.core.*

Python/PyTorch host code over hand-written sm_100a CUDA (``csrc/`` -> ``libtdyn_b200.so``, C ABI in
``include/tdyn_b200.h``).  No CPU fallback, no Triton, no multi-backend dispatch.
"""
from . import distributed  # noqa: F401
from .core import DEFunc, DEFuncBase, NeuralODE, ODEProblem
from .numerics import (SOLVER_DICT, DormandPrince45, Euler, Midpoint, RungeKutta4, Tsitouras45, odeint, str_to_solver)

__version__ = "0.1.0"
__all__ = ["odeint", "NeuralODE", "ODEProblem", "DEFunc", "DEFuncBase", "Euler", "Midpoint", "RungeKutta4", "DormandPrince45",
           "Tsitouras45", "SOLVER_DICT", "str_to_solver", "distributed"]
