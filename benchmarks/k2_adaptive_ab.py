#!/usr/bin/env python
"""A/B: adaptive (dopri5 / tsit5) forward solve of the 32-256-256-32 class -- ONE cooperative launch of the tensor-core
stage kernel (tdyn_mlp_tc_solve_adaptive) against the host-driven loop over the same kernel (one launch per stage,
tdyn_rk_finish, one host synchronisation per attempted step).  Alternating, CUDA events, same box.

    python benchmarks/k2_adaptive_ab.py [--rows 1048576] [--reps 5] [--gain 1.0] [--tol 1e-4]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torchdyn_b200 import fused  # noqa: E402
from torchdyn_b200.numerics import odeint  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--gain", type=float, default=1.0)
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--solver", default="dopri5")
    ap.add_argument("--t1", type=float, default=1.0)
    ap.add_argument("--only", default="", help="one_launch | host_loop: run only that mode (profiling)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(32, 256), nn.Tanh(), nn.Linear(256, 256), nn.Tanh(), nn.Linear(256, 32)).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(a.gain)
    x = torch.randn(a.rows, 32, device=dev)
    fld = fused.MLPField(net)
    fused.ADAPTIVE_TC_MAX_ROWS = None        # A/B at every size
    ts = torch.tensor([0.0, a.t1])
    res = {True: [], False: []}
    nfe = {}
    sols = {}
    for rep in range(a.reps + 2):
        for flag in ((True,) if a.only == "one_launch" else (False,) if a.only == "host_loop" else (True, False)):
            fused.ADAPTIVE_TC = flag
            fld.nfe = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            with torch.no_grad():
                _, sol = odeint(fld, x, ts, a.solver, atol=a.tol, rtol=a.tol)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                res[flag].append(e0.elapsed_time(e1))
            nfe[flag] = fld.nfe
            sols[flag] = sol[-1]
            path = fused.last_used()
            if rep == 0:
                print(flag, path, file=sys.stderr)
    fused.ADAPTIVE_TC = True
    flops = 2.0 * (32 * 256 + 256 * 256 + 256 * 32)
    out = {"rows": a.rows, "solver": a.solver, "tol": a.tol, "gain": a.gain, "nfe": nfe,
           "bit_identical": bool(torch.equal(sols[True], sols[False])) if len(sols) == 2 else None}
    for flag, name in ((True, "one_launch"), (False, "host_loop")):
        if not res[flag]:
            continue
        ms = sorted(res[flag])[len(res[flag]) // 2]
        out[name] = {"ms": round(ms, 3), "sample_nfe_per_s": a.rows * nfe[flag] / ms * 1e3,
                     "alg_tflops": flops * a.rows * nfe[flag] / ms * 1e-9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
