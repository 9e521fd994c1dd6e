"""``standardize_vf_call_signature`` (reference torchdyn/core/utils.py:19-35): every vector field handed to ODEProblem /
NeuralODE ends up callable as ``(t, x)``."""
from inspect import getfullargspec

import torch.nn as nn

from .defunc import DEFunc, DEFuncBase

_NOTICE = ("Your vector field callable ({kind}) should have both time `t` and state `x` as arguments, "
           "we've wrapped it for you.")


def _takes_time(vf) -> bool:
    """True if the callable's own signature names a ``t`` argument (a module is judged by its ``forward``)."""
    target = vf.forward if isinstance(vf, nn.Module) else vf
    return 't' in getfullargspec(target).args


def standardize_vf_call_signature(vector_field, order=1, defunc_wrap=False):
    is_module = isinstance(vector_field, nn.Module)
    timed = _takes_time(vector_field)
    if not timed:
        print(_NOTICE.format(kind="nn.Module" if is_module else "lambda"))
    if not (is_module and timed):            # a module that already takes (t, x) is used as it is
        vector_field = DEFuncBase(vector_field, has_time_arg=timed)
    return DEFunc(vector_field, order) if defunc_wrap else vector_field
