"""Continuous normalizing flow vector field ``[-tr(df/dx), f(x)]`` (reference torchdyn/models/cnf.py:19-65).

A CNF is a *workload* of the integration path (BASELINE config 4): a vector field whose evaluation itself uses autograd
(the Jacobian trace), so it runs through the generic CUDA path (stage-combine / finish kernels, f via torch) and its
adjoint needs second-order autograd, exactly as in the reference."""
from typing import Callable, Union

import torch
import torch.nn as nn
from torch.autograd import grad


def autograd_trace(x_out, x_in, **kwargs):
    """Exact Jacobian trace with one autograd call per state dimension."""
    tr = 0.
    for i in range(x_in.shape[1]):
        tr += grad(x_out[:, i].sum(), x_in, allow_unused=False, create_graph=True)[0][:, i]
    return tr


def hutch_trace(x_out, x_in, noise=None, **kwargs):
    """Hutchinson estimator eps^T J eps with a single vector-Jacobian product."""
    vjp = grad(x_out, x_in, noise, create_graph=True)[0]
    return torch.einsum('bi,bi->b', vjp, noise)


REQUIRES_NOISE = [hutch_trace]


class CNF(nn.Module):
    def __init__(self, net: nn.Module, trace_estimator: Union[Callable, None] = None, noise_dist=None, order=1):
        """``net``: data vector field; ``trace_estimator``: autograd_trace (default) or hutch_trace (needs
        ``noise_dist`` with a ``.sample`` method; NeuralODE samples ``noise`` once per forward)."""
        super().__init__()
        self.net, self.order = net, order
        self.trace_estimator = trace_estimator if trace_estimator is not None else autograd_trace
        self.noise_dist, self.noise = noise_dist, None
        if self.trace_estimator in REQUIRES_NOISE:
            assert self.noise_dist is not None, 'This type of trace estimator requires specification of a noise distribution'

    def forward(self, x):
        with torch.set_grad_enabled(True):
            x_in = x[:, 1:].requires_grad_(True)          # column 0 carries the accumulated divergence
            x_out = self.net(x_in)
            tr = self.trace_estimator(x_out, x_in, noise=self.noise)
        return torch.cat([-tr[:, None], x_out], 1) + 0 * x   # `+ 0*x` ties x[:, 0] into the graph, as in the reference
