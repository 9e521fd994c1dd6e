"""``standardize_vf_call_signature`` (reference torchdyn/core/utils.py:19-35)."""
from inspect import getfullargspec

import torch.nn as nn

from .defunc import DEFunc, DEFuncBase


def standardize_vf_call_signature(vector_field, order=1, defunc_wrap=False):
    "Make callables / nn.Modules handed to ODEProblem and NeuralODE callable as (t, x)."
    if issubclass(type(vector_field), nn.Module):
        if 't' not in getfullargspec(vector_field.forward).args:
            print("Your vector field callable (nn.Module) should have both time `t` and state `x` as arguments, "
                  "we've wrapped it for you.")
            vector_field = DEFuncBase(vector_field, has_time_arg=False)
    else:
        # plain functions / lambdas: inspect the function itself
        if 't' not in getfullargspec(vector_field).args:
            print("Your vector field callable (lambda) should have both time `t` and state `x` as arguments, "
                  "we've wrapped it for you.")
            vector_field = DEFuncBase(vector_field, has_time_arg=False)
        else:
            vector_field = DEFuncBase(vector_field, has_time_arg=True)
    return DEFunc(vector_field, order) if defunc_wrap else vector_field
