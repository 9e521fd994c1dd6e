#!/usr/bin/env python
"""Diagnostic: run fixtures on the GPU and print the reference's and this library's step traces side by side.
    python benchmarks/diag_fixture.py tests/golden/act_w1024_tsit5_interp.pt [...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_helpers import run_neuralode_case  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
for path in sys.argv[1:]:
    fx = torch.load(path, weights_only=False)
    out = run_neuralode_case(fx, "cuda:0")
    ref = fx["ref"]
    print("==", os.path.basename(path))
    for r, g in zip(ref["traces"], out["traces"]):
        print("  ref n=%d ratios %s" % (len(r["t"]), ["%.3g" % v for v in r["ratio"]]))
        print("  gpu n=%d ratios %s" % (len(g["t"]), ["%.3g" % v for v in g["ratio"]]))
        n = min(len(r["t"]), len(g["t"]))
        print("      max |dt-dt_ref|/dt_ref over common attempts: %.2e" % max(abs(a - b) / abs(b) for a, b in zip(g["dt"][:n], r["dt"][:n])))

    def rel(a, b):
        return float((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max())
    print("  state %.2e  gx %.2e  gp %s" % (rel(out["sol"][-1], ref["sol_last"]), rel(out["x"].grad, ref["gx"]),
                                             ["%.2e" % rel(p.grad, g) for p, g in zip(out["net"].parameters(), ref["gp"])]))
