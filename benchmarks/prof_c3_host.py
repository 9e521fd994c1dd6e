"""Host-side profile of a config-3 training iteration at 1/8 of the batch (the per-GPU share of the 8-GPU strong-scaling
point, where the launches -- not the kernels -- bound the iteration)."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import run_configs as rc  # noqa: E402
import torchdyn_b200  # noqa: E402
from torchdyn_b200 import _kernels  # noqa: E402

dev = torch.device("cuda:0")
net, X, t_span, kw, loss_fn, name = rc.build(3, dev, 0.125)
net = net.to(dev)
model = torchdyn_b200.NeuralODE(net, **kw)
x = X.to(dev)


def it():
    for p in net.parameters():
        p.grad = None
    xx = x.clone().requires_grad_(True)
    _, sol = model(xx, t_span)
    loss_fn(sol).backward()


for _ in range(5):
    it()
torch.cuda.synchronize()
l0 = _kernels.launches()
t0 = time.perf_counter()
for _ in range(20):
    it()
torch.cuda.synchronize()
print("ms/iter", 1e3 * (time.perf_counter() - t0) / 20, "launches/iter", (_kernels.launches() - l0) / 20)
timer = _kernels.KernelTimer()
_kernels.set_timer(timer)
it()
_kernels.set_timer(None)
print("kernel ms/iter", sum(v["ms"] for v in timer.summary().values()), {k: (v["launches"], round(v["ms"], 2)) for k, v in timer.summary().items()})
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    it()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(30)
print(s.getvalue()[:6000])
