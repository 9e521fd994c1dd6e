"""Augmenter / DepthCat / DataControl (reference torchdyn/nn/node_layers.py:16-82).  Torch modules, no kernels: they
run before the solve (Augmenter) or inside a vector field that then takes the generic path (DepthCat, DataControl)."""
import torch
import torch.nn as nn


class Augmenter(nn.Module):
    """Adds ``augment_dims`` channels along ``augment_idx``: zeros, or ``augment_func(x)`` when given; placed before
    (``order='first'``) or after the data channels."""

    def __init__(self, augment_idx: int = 1, augment_dims: int = 5, augment_func=None, order='first'):
        super().__init__()
        self.augment_dims, self.augment_idx, self.augment_func = augment_dims, augment_idx, augment_func
        self.order = order

    def forward(self, x: torch.Tensor):
        if self.augment_func:
            extra = self.augment_func(x).to(x)
        else:
            shape = list(x.shape)
            shape[self.augment_idx] = self.augment_dims
            extra = torch.zeros(shape).to(x)
        parts = [extra, x] if self.order == 'first' else [x, extra]
        return torch.cat(parts, self.augment_idx)


class DepthCat(nn.Module):
    """Concatenates the depth variable ``t`` (injected by DEFunc before every evaluation) along ``idx_cat``."""

    def __init__(self, idx_cat=1):
        super().__init__()
        self.idx_cat, self.t = idx_cat, None

    def forward(self, x):
        shape = list(x.shape)
        shape[self.idx_cat] = 1
        return torch.cat([x, self.t * torch.ones(shape).to(x)], self.idx_cat).to(x)


class DataControl(nn.Module):
    """Concatenates the (fixed) initial condition assigned by NeuralODE._prep_integration to the current state."""

    def __init__(self):
        super().__init__()
        self._control = None

    def forward(self, x):
        return torch.cat([x, self._control], 1).to(x)
