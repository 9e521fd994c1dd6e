import os, sys, cProfile, pstats, io, time
sys.path.insert(0, '/root/repo')
import torch
from benchmarks import run_configs as rc
import torchdyn_b200
dev = torch.device("cuda:0")
net, X, t_span, kw, loss_fn, name = rc.build(1, dev)
net = net.to(dev); model = torchdyn_b200.NeuralODE(net, **kw); x = X.to(dev)
def it():
    for p in net.parameters(): p.grad = None
    xx = x.clone().requires_grad_(True)
    _, sol = model(xx, t_span)
    loss_fn(sol).backward()
for _ in range(20): it()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(200): it()
torch.cuda.synchronize()
print("ms/iter", 1e3*(time.perf_counter()-t0)/200)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): it()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40); print(s.getvalue()[:9000])
