import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torchdyn_b200
from torchdyn_b200 import fused
from torchdyn_b200.numerics import odeint
from conftest import build_mlp
res = []
for d in (0, 1, 0):
    dev = f"cuda:{d}"
    torch.manual_seed(0)
    net = build_mlp([32, 256, 256, 32], "tanh").to(dev)
    x = torch.randn(1000, 32, device=dev)
    with torch.cuda.device(d), torch.no_grad():
        _, s = odeint(fused.MLPField(net), x, torch.linspace(0, 1, 5), "rk4")
        net2 = build_mlp([64, 256, 64], "tanh").to(dev)
        x2 = torch.randn(300, 64, device=dev)
        _, s2 = odeint(fused.MLPField(net2), x2, torch.linspace(0, 1, 3), "dopri5")
    res.append((s[-1].cpu(), s2[-1].cpu(), fused.last_used()))
    print(dev, float(s[-1].abs().mean()), float(s2[-1].abs().mean()), fused.last_used())
assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][0], res[2][0])
print("two-device ok")
