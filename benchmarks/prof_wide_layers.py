"""ncu target: one launch each of the c3-shaped layer kernels (after a warm-up launch)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdyn_b200 import _kernels
from torchdyn_b200._lib import ACT_TANH
dev = torch.device("cuda:0")
kern = _kernels.get(torch.empty(1, device=dev))
rows = 262144
for k_in, n_out in ((128, 1024), (1024, 128)):
    w = torch.randn(n_out, k_in, device=dev) / k_in ** 0.5
    b = torch.randn(n_out, device=dev)
    x = torch.randn(rows, k_in, device=dev)
    g = torch.randn(rows, n_out, device=dev)
    out = torch.empty(rows, n_out, device=dev)
    gin = torch.empty(rows, k_in, device=dev)
    img, img_t = kern.lin_tc_prepare(w, False), kern.lin_tc_prepare(w, True)
    dw, db = torch.empty_like(w), torch.empty_like(b)
    ws = kern.wgrad_tc_workspace(n_out, k_in, rows)
    for _ in range(2):
        kern.lin_tc_apply(img, n_out, k_in, x, b, None, out, ACT_TANH, 1, 1.0, rows)
        kern.lin_tc_apply(img_t, k_in, n_out, g, None, x, gin, ACT_TANH, 2, 1.0, rows)
        kern.wgrad_tc(g, n_out, x, k_in, dw, db, -1.0, ws, rows)
    torch.cuda.synchronize()
