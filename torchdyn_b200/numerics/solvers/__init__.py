from .ode import (BUILTIN_TYPES, SOLVER_DICT, AsynchronousLeapfrog, DormandPrince45, Euler, Midpoint, RungeKutta4, Tsitouras45, str_to_solver)
from .hyper import HyperEuler, HyperMidpoint, HyperRungeKutta4
from .templates import DiffEqSolver

SolverTemplate = DiffEqSolver

__all__ = ["DiffEqSolver", "SolverTemplate", "Euler", "Midpoint", "RungeKutta4", "DormandPrince45", "Tsitouras45", "AsynchronousLeapfrog",
           "SOLVER_DICT", "str_to_solver", "BUILTIN_TYPES", "HyperEuler", "HyperMidpoint", "HyperRungeKutta4"]
