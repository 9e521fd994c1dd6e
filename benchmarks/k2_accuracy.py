import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from torchdyn_b200 import _kernels, fused
from torchdyn_b200._lib import MODE_INNER
from conftest import build_mlp
dev="cuda:0"
kern=_kernels.get(torch.empty(1,device=dev))
for act in ("tanh","softplus","relu"):
    for seed in (8, 9):
        torch.manual_seed(seed)
        net=build_mlp([32,256,256,32],act).to(dev)
        fm=fused.FusedMLP(net); assert fm.ready(kern)
        x=torch.randn(8192,32,device=dev)
        out=torch.empty_like(x)
        kern.mlp_tc_stage(fm._desc, fm._img, out, x, [], (), MODE_INNER, 0.0, None, (), fm.flops_per_row)
        net64=build_mlp([32,256,256,32],act,[p.detach() for p in net.parameters()],dev).double()
        with torch.no_grad(): want=net64(x.double())
            # fp32 torch reference too
        with torch.no_grad(): f32=net(x)
        sc=float(want.abs().max())
        print(act, seed, "max err / scale: kernel %.2e   torch fp32 %.2e" % (float((out.double()-want).abs().max())/sc, float((f32.double()-want).abs().max())/sc))
