#!/usr/bin/env python
"""Per-launch timing of the layer-wise tensor-core kernels (tdyn_lin_tc_apply, tdyn_wgrad_tc) at the layer shapes of
BASELINE config 3 (128-1024-128, 262 144 rows) and of the 32-256-256-32 class (2^20 rows).  One JSON line per shape:
algorithmic TFLOP/s = 2 * n_out * k_in * rows / time (each FLOP is issued three times as TF32).

    python benchmarks/wide_tc_layers.py [--iters 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdyn_b200 import _kernels  # noqa: E402
from torchdyn_b200._lib import ACT_TANH  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    kern = _kernels.get(torch.empty(1, device=dev))
    cases = [("c3", 262144, 128, 1024), ("c3", 262144, 1024, 128), ("k2", 1 << 20, 32, 256), ("k2", 1 << 20, 256, 256),
             ("k2", 1 << 20, 256, 32)]
    for tag, rows, k_in, n_out in cases:
        torch.manual_seed(0)
        w = torch.randn(n_out, k_in, device=dev) / k_in ** 0.5
        b = torch.randn(n_out, device=dev)
        x = torch.randn(rows, k_in, device=dev)
        g = torch.randn(rows, n_out, device=dev)
        out = torch.empty(rows, n_out, device=dev)
        gin = torch.empty(rows, k_in, device=dev)
        img, img_t = kern.lin_tc_prepare(w, False), kern.lin_tc_prepare(w, True)
        dw, db = torch.empty_like(w), torch.empty_like(b)
        ws = kern.wgrad_tc_workspace(n_out, k_in, rows)
        flop = 2.0 * n_out * k_in * rows
        runs = {
            "forward+act": lambda: kern.lin_tc_apply(img, n_out, k_in, x, b, None, out, ACT_TANH, 1, 1.0, rows),
            "forward linear": lambda: kern.lin_tc_apply(img, n_out, k_in, x, b, None, out, ACT_TANH, 0, 1.0, rows),
            "dgrad*act'": lambda: kern.lin_tc_apply(img_t, k_in, n_out, g, None, x, gin, ACT_TANH, 2, 1.0, rows),
            "wgrad": lambda: kern.wgrad_tc(g, n_out, x, k_in, dw, db, -1.0, ws, rows),
        }
        for name, fn in runs.items():
            ms = timed(fn, a.iters)
            print(json.dumps({"shape": f"{tag}: {k_in}->{n_out}, rows {rows}", "kernel": name, "ms": round(ms, 4),
                              "alg_tflops": round(flop / ms / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
