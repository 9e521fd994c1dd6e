from .odeint import EventState, odeint, odeint_hybrid, odeint_symplectic
from .solvers.ode import (SOLVER_DICT, AsynchronousLeapfrog, DormandPrince45, Euler, Midpoint, RungeKutta4, Tsitouras45, str_to_solver)
from .solvers.templates import DiffEqSolver
from .interpolators import FourthOrder, INTERP_DICT, str_to_interp
from .utils import adapt_step, hairer_norm, init_step

__all__ = ['odeint', 'odeint_symplectic', 'odeint_hybrid', 'EventState', 'AsynchronousLeapfrog', 'Euler', 'Midpoint', 'RungeKutta4', 'DormandPrince45', 'Tsitouras45', 'DiffEqSolver', 'SOLVER_DICT',
           'str_to_solver', 'FourthOrder', 'INTERP_DICT', 'str_to_interp', 'hairer_norm', 'init_step', 'adapt_step']
