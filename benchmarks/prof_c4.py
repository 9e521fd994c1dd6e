#!/usr/bin/env python
"""Two training iterations of config 4 (CNF / Hutchinson, 65 536 rows): the command profiled for the FFMA kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import run_configs as rc  # noqa: E402
import torchdyn_b200  # noqa: E402

dev = torch.device("cuda:0")
net, X, t_span, kw, loss_fn, name = rc.build(4, dev)
net = net.to(dev)
model = torchdyn_b200.NeuralODE(net, **kw)
x = X.to(dev)
for _ in range(2):
    for p in net.parameters():
        p.grad = None
    _, sol = model(x.clone().requires_grad_(True), t_span)
    loss_fn(sol).backward()
torch.cuda.synchronize()
print("ok", float(model.vf.nfe))
