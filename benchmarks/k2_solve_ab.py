#!/usr/bin/env python
"""A/B: c2 (rk4 x100, 32-256-256-32, 2^20 rows) with one launch per solve (tdyn_mlp_tc_solve_fixed) vs one launch per stage
(tdyn_mlp_tc_stage), alternating in one process, with the SM clock sampled during each leg."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from torchdyn_b200 import fused  # noqa: E402
from torchdyn_b200.numerics import odeint  # noqa: E402

dev = torch.device("cuda:0")
net = bench.make_net(dev)
x = torch.randn(1 << 20, 32, device=dev)
ts = torch.linspace(0, 1, 101)
fld = fused.MLPField(net)


def leg(flag, n=4):
    fused.SOLVE_KERNEL = flag
    with torch.no_grad():
        odeint(fld, x, ts, "rk4", save_at=ts[-1:])
        torch.cuda.synchronize()
        with bench.ClockSampler(0) as cs:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                odeint(fld, x, ts, "rk4", save_at=ts[-1:])
            e1.record()
            torch.cuda.synchronize()
    return {"one_launch_per_solve": flag, "ms_per_solve": e0.elapsed_time(e1) / n, "clocks": cs.result}


for rep in range(2):
    for flag in (False, True):
        print(json.dumps(leg(flag)), flush=True)
