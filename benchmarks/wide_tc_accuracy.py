#!/usr/bin/env python
"""Measured error of the layer-wise tensor-core kernels against fp64 (max |err| / max |reference|), next to torch fp32."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdyn_b200 import _kernels
from torchdyn_b200._lib import ACT_TANH
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
kern = _kernels.get(torch.empty(1, device=dev))
for n_out, k_in, rows in ((1024, 128, 4096), (128, 1024, 4096), (256, 256, 4096), (128, 128, 65536)):
    torch.manual_seed(0)
    w = torch.randn(n_out, k_in, device=dev) / k_in ** 0.5
    b = torch.randn(n_out, device=dev)
    x = torch.randn(rows, k_in, device=dev)
    g = torch.randn(rows, n_out, device=dev)
    out = torch.empty(rows, n_out, device=dev)
    kern.lin_tc_apply(kern.lin_tc_prepare(w, False), n_out, k_in, x, b, None, out, ACT_TANH, 0, 1.0, rows)
    want = x.double() @ w.double().T + b.double()
    ref32 = x @ w.T + b
    dw, db = torch.empty_like(w), torch.empty_like(b)
    kern.wgrad_tc(g, n_out, x, k_in, dw, db, 1.0, kern.wgrad_tc_workspace(n_out, k_in, rows), rows)
    want_w = g.double().T @ x.double()
    ref_w = g.T @ x
    rel = lambda a, r: float((a.double() - r).abs().max() / r.abs().max())
    print(json.dumps({"layer": f"{k_in}->{n_out}", "rows": rows, "forward_err": rel(out, want), "forward_err_torch_fp32": rel(ref32, want),
                      "wgrad_err": rel(dw, want_w), "wgrad_err_torch_fp32": rel(ref_w, want_w)}))
