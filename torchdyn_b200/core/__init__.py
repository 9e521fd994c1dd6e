from .defunc import DEFunc, DEFuncBase
from .neuralde import NeuralODE
from .problems import ODEProblem

__all__ = ["ODEProblem", "NeuralODE", "DEFunc", "DEFuncBase"]
