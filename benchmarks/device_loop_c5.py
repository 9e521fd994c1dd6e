#!/usr/bin/env python
"""Config 5 forward (Augmenter(10) + Conv2d(11,11,3) + Tanh on 28x28, batch 4096, dopri5 1e-3, no_grad): host-driven loop
(one 8-byte read per attempted step) vs the device-side loop (one CUDA-graph launch per solve, one read per solve)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks import run_configs as rc  # noqa: E402
from torchdyn_b200.numerics import graph_loop, odeint  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
om = sys.modules["torchdyn_b200.numerics.odeint"]
for B in (4096, 256):
    net, X, t_span, kw, _, name = rc.build(5, dev, B / 4096)
    net = net.to(dev)
    x = X.to(dev)
    f = lambda t, y: net(y)      # noqa: E731
    out = {"workload": name.replace("backsolve adjoint", "forward only"), "batch": B}
    for label, flag, cache in (("host_loop", False, False), ("device_loop", True, False), ("device_loop_cached_graph", True, True)):
        graph_loop.ENABLED, graph_loop.CACHE = flag, cache
        om.TRACE_SINK = tr = []
        with torch.no_grad():
            odeint(f, x, t_span, "dopri5", atol=1e-3, rtol=1e-3)
            om.TRACE_SINK = None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                _, sol = odeint(f, x, t_span, "dopri5", atol=1e-3, rtol=1e-3)
            torch.cuda.synchronize()
        out[label] = {"ms_per_solve": 1e3 * (time.perf_counter() - t0) / 5, "attempts": len(tr[0]["t"]), "status": graph_loop.last_status,
                      "checksum": float(sol[-1].double().sum())}
    graph_loop.ENABLED = graph_loop.CACHE = False
    print(json.dumps(out), flush=True)
