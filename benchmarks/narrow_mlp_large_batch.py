#!/usr/bin/env python
"""Narrow MLP (2-64-64-64-2 softplus), 65 536 rows, dopri5 + backsolve adjoint through the persistent kernels: the
large-batch regime of the FFMA path (FP32-issue bound rather than latency bound)."""
import json
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdyn_b200  # noqa: E402
from torchdyn_b200 import fused  # noqa: E402


def main(iters=10, rows=65536):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(2, 64), nn.Softplus(), nn.Linear(64, 64), nn.Softplus(), nn.Linear(64, 64), nn.Softplus(),
                        nn.Linear(64, 2)).to(dev)
    model = torchdyn_b200.NeuralODE(net, solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint")
    x = torch.randn(rows, 2, device=dev)

    def it():
        for p in net.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        _, sol = model(xx, torch.linspace(0, 1, 2))
        sol[-1].pow(2).mean().backward()

    for _ in range(3):
        it()
    torch.cuda.synchronize()
    n0, t0 = model.vf.nfe, time.perf_counter()
    for _ in range(iters):
        it()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    nfe = (model.vf.nfe - n0) / iters
    flops = 2 * (2 * 64 + 64 * 64 * 2 + 64 * 2)
    print(json.dumps({"config": f"MLP 2-64-64-64-2 softplus, {rows} rows, dopri5 + backsolve adjoint", "ms_per_iter": 1e3 * dt,
                      "nfe_per_iter": nfe, "sample_nfe_per_s": rows * nfe / dt, "path": fused.last_used(),
                      "fwd_flop_per_sample_nfe": flops}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
