import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from torchdyn_b200 import _kernels, fused
from torchdyn_b200._lib import MODE_INNER
from conftest import build_mlp
dev="cuda:0"
kern=_kernels.get(torch.empty(1,device=dev))
x=torch.randn(1<<20,32,device=dev); out=torch.empty_like(x)
for act in ("tanh","softplus","relu"):
    net=build_mlp([32,256,256,32],act).to(dev)
    fm=fused.FusedMLP(net); assert fm.ready(kern)
    f=lambda: kern.mlp_tc_stage(fm._desc, fm._img, out, x, [], (), MODE_INNER, 0.0, None, (), fm.flops_per_row)
    for _ in range(20): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): f()
    e1.record(); torch.cuda.synchronize()
    print(act, "ms/launch %.4f" % (e0.elapsed_time(e1)/200))
