#!/usr/bin/env python
"""Secondary measurements: every BASELINE.json config through the public API on one B200 (the headline, config 2, is
bench.py).  One JSON line per config: sample-NFE/s = batch * NFE / time for one training iteration (forward + loss +
backward where the config has a backward), NFE as counted by ``model.vf.nfe`` (reference core/defunc.py:61).

    python benchmarks/run_configs.py [--configs 1,3,5] [--iters 5] [--cpu]      # --cpu: also time the CPU oracle port
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def moons(n, noise=0.4, seed=0):
    """Two interleaved half circles + uniform noise (same construction as the reference's ToyDataset moons,
    datasets/static_datasets.py:71-95; regenerated here because the reference tree does not exist on the GPU box)."""
    rng = np.random.RandomState(seed)
    n_out, n_in = n // 2, n - n // 2
    outer = np.stack([np.cos(np.linspace(0, np.pi, n_out)), np.sin(np.linspace(0, np.pi, n_out))], 1)
    inner = np.stack([1 - np.cos(np.linspace(0, np.pi, n_in)), 1 - np.sin(np.linspace(0, np.pi, n_in)) - .5], 1)
    X = np.vstack([outer, inner]) + noise * rng.rand(n, 2)
    y = np.hstack([np.zeros(n_out), np.ones(n_in)])
    return torch.tensor(X, dtype=torch.float32), torch.tensor(y, dtype=torch.long)


def mlp(dims, act=nn.Tanh):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(act())
    return nn.Sequential(*layers)


class Augment(nn.Module):
    """zero-augmentation of the channel dim (what torchdyn.nn.Augmenter(augment_dims=k) does by default)"""

    def __init__(self, dims):
        super().__init__()
        self.dims = dims

    def forward(self, x):
        z = torch.zeros(x.shape[0], self.dims, *x.shape[2:], device=x.device, dtype=x.dtype)
        return torch.cat([x, z], 1)


def build(cfg, device, scale=1.0):
    torch.manual_seed(0)
    np.random.seed(0)
    if cfg == 1:
        X, y = moons(1024)
        net = mlp([2, 64, 2])
        kw = dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint")
        loss = lambda sol, y=y.to(device): nn.CrossEntropyLoss()(sol[-1], y)   # noqa: E731
        return net, X, torch.linspace(0, 1, 2), kw, loss, "c1 moons 1024x2, MLP 2-64-2, dopri5 1e-4, backsolve adjoint"
    if cfg == 2:
        # the north-star headline on the config-2 network: dopri5 forward + backsolve adjoint (not a BASELINE config by
        # itself: configs[1] is rk4 forward only = bench.py); batch 2^18 so that the adjoint state fits comfortably
        B = int(262144 * scale)
        net = mlp([32, 256, 256, 32])
        X = torch.randn(B, 32)
        kw = dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint", atol_adjoint=1e-4, rtol_adjoint=1e-4)
        return net, X, torch.linspace(0, 1, 2), kw, (lambda sol: sol[-1].pow(2).mean()), \
            f"c2a dopri5 1e-4 forward + backsolve adjoint, MLP 32-256-256-32, batch {B}"
    if cfg == 3:
        B = int(262144 * scale)
        net = mlp([128, 1024, 128])
        X = torch.randn(B, 128)
        kw = dict(solver="tsit5", sensitivity="interpolated_adjoint")
        return net, X, torch.linspace(0, 1, 2), kw, (lambda sol: sol[-1].pow(2).mean()), \
            f"c3 tsit5 + interpolated adjoint, MLP 128-1024-128, batch {B}"
    if cfg == 4:
        from torchdyn_b200.models import CNF, hutch_trace
        from torchdyn_b200.nn import Augmenter
        B = int(65536 * scale)
        net = mlp([2, 64, 64, 64, 2], nn.Softplus)
        noise_dist = torch.distributions.MultivariateNormal(torch.zeros(2, device=device), torch.eye(2, device=device))
        cnf = CNF(net, trace_estimator=hutch_trace, noise_dist=noise_dist)
        X = Augmenter(1, 1)(torch.randn(B, 2))
        kw = dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint")

        def loss(sol, nd=noise_dist):
            return -(nd.log_prob(sol[-1][:, 1:]) - sol[-1][:, 0]).mean()
        return cnf, X, torch.linspace(0, 1, 2), kw, loss, f"c4 CNF hutchinson, MLP 2-64-64-64-2 softplus, dopri5 1e-4, adjoint, batch {B}"
    if cfg == 5:
        B = int(4096 * scale)
        net = nn.Sequential(nn.Conv2d(11, 11, 3, padding=1), nn.Tanh())
        X = Augment(10)(torch.randn(B, 1, 28, 28))
        kw = dict(solver="dopri5", atol=1e-3, rtol=1e-3, sensitivity="adjoint")
        return net, X, torch.linspace(0, 1, 2), kw, (lambda sol: sol[-1].pow(2).mean()), \
            f"c5 conv 11ch 28x28 (augmented), dopri5 1e-3, backsolve adjoint, batch {B}"
    raise ValueError(cfg)


def run_gpu(cfg, iters, scale):
    import torchdyn_b200
    from torchdyn_b200 import _kernels
    dev = torch.device("cuda:0")
    net, X, t_span, kw, loss_fn, name = build(cfg, dev, scale)
    net = net.to(dev)
    model = torchdyn_b200.NeuralODE(net, **kw)
    x = X.to(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def it():
        for p in net.parameters():
            p.grad = None
        xx = x.clone().requires_grad_(True)
        _, sol = model(xx, t_span)
        loss_fn(sol).backward()

    for _ in range(3):
        it()
    torch.cuda.synchronize()
    timer = _kernels.KernelTimer()
    _kernels.set_timer(timer)
    it()
    _kernels.set_timer(None)
    per_kernel = {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in timer.summary().items()}
    n0, l0 = model.vf.nfe, _kernels.launches()
    t0 = time.perf_counter()
    for _ in range(iters):
        it()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    nfe = (model.vf.nfe - n0) / iters
    return dict(config=name, impl="b200", batch=x.shape[0], nfe_per_iter=nfe, ms_per_iter=1e3 * dt,
                sample_nfe_per_s=x.shape[0] * nfe / dt, own_kernel_launches_per_iter=(_kernels.launches() - l0) / iters,
                own_kernel_ms_per_iter=per_kernel)


def run_cpu(cfg, iters, scale):
    from oracle import erk_oracle as eo
    torch.set_num_threads(os.cpu_count() or 1)
    net, X, t_span, kw, loss_fn, name = build(cfg, torch.device("cpu"), scale)
    if getattr(net, "noise_dist", None) is not None:      # CNF: NeuralODE samples the Hutchinson noise once per forward
        net.noise = net.noise_dist.sample((X.shape[0],))
    model = eo.OracleNeuralODE(net, **kw)

    def it():
        for p in net.parameters():
            p.grad = None
        xx = X.clone().requires_grad_(True)
        _, sol = model(xx, t_span)
        loss_fn(sol).backward()

    it()
    n0 = model.nfe
    t0 = time.perf_counter()
    for _ in range(iters):
        it()
    dt = (time.perf_counter() - t0) / iters
    nfe = (model.nfe - n0) / iters
    return dict(config=name, impl="cpu oracle port", cores=os.cpu_count(), batch=X.shape[0], nfe_per_iter=nfe,
                ms_per_iter=1e3 * dt, sample_nfe_per_s=X.shape[0] * nfe / dt)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--scale", type=float, default=1.0, help="batch scale for configs 3 and 5")
    ap.add_argument("--cpu-scale", type=float, default=1 / 32)
    a = ap.parse_args()
    for c in [int(v) for v in a.configs.split(",")]:
        print(json.dumps(run_gpu(c, a.iters, a.scale)), flush=True)
        if a.cpu:
            print(json.dumps(run_cpu(c, max(1, a.iters // 2), 1.0 if c == 1 else (1 / 8 if c == 4 else a.cpu_scale))), flush=True)
