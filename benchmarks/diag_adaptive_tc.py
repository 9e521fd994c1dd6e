"""Diagnostic: the wide active functional fixtures on (a) the one-launch adaptive tensor-core integrator, (b) the host-driven
loop over the same stage kernel, (c) the generic path with f as a torch fp32 call -- error against the executed reference
per output time, and (a) against (b)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_mlp, load_golden  # noqa: E402
from parity_helpers import run_functional_case  # noqa: E402
from torchdyn_b200 import fused  # noqa: E402

dev = torch.device("cuda:0")
for name in ("act_w256x2_tsit5_fn", "act_w256x2_dopri5_fn"):
    fx = load_golden(name)
    net = build_mlp(fx["dims"], fx["act"], fx["params"], dev)
    ref = fx["ref"]["sol"].double()
    res = {}
    for label, flag, f in (("a", True, fused.MLPField(net)), ("b", False, fused.MLPField(net)), ("c", True, None)):
        fused.ADAPTIVE_TC = flag
        t_eval, sol, tr = run_functional_case(fx, dev, f=f)
        fused.ADAPTIVE_TC = True
        res[label] = sol.double().cpu()
        ok = sol.shape == ref.shape
        e = [(float((res[label][i] - ref[i]).abs().max()), float(ref[i].abs().max())) for i in range(min(len(ref), len(sol)))]
        print(name, label, fused.last_used(), "shape ok" if ok else "SHAPE", "attempts", len(tr[0]["t"]))
        print("   max|err| / max|ref| per output:", " ".join(f"{a:.2e}/{b:.1f}" for a, b in e))
        print("   ratios", [round(r, 4) for r in tr[0]["ratio"]])
    print("   a vs b max diff per output:", [float((res["a"][i] - res["b"][i]).abs().max()) for i in range(len(res["a"]))])
