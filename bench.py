#!/usr/bin/env python
"""Headline benchmark of the batched ODE-integration hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "c2"): rk4 fixed-step, 100 steps over t in [0, 1], MLP vector field
Linear(32,256)-Tanh-Linear(256,256)-Tanh-Linear(256,32) (random init, seed 0), synthetic state x ~ N(0,1) of
2^20 rows per GPU (weak scaling: every rank integrates its own 2^20-row shard, no data-path collective -- rk4 is
fixed-step), forward only, only the final state saved.  One "step" = one full solve = 400 vector-field
evaluations per sample.  metric = sample-NFE/s = rows * 400 / time, summed over ranks.

Lines printed (rank 0): ONE JSON object with value (inputs resident in HBM), e2e (host pinned buffers in, host
result out, copies inside the timed region), roofline (dominant kernel of this library, CUDA-event timed inside
the timed region), cpu_baseline (the CPU oracle port on a bounded sub-batch), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (32, 256, 256, 32)
N_STEPS_RK = 100
NFE_PER_SOLVE = 4 * N_STEPS_RK
ROWS_PER_GPU = 1 << 20
CPU_SAMPLE_ROWS = 16384
METRIC = "sample-NFE/sec"
FLOPS_PER_SAMPLE_NFE = 2 * sum(a * b for a, b in zip(DIMS[:-1], DIMS[1:]))     # 163 840


def make_net(device):
    torch.manual_seed(0)
    layers = []
    for i in range(len(DIMS) - 1):
        layers.append(nn.Linear(DIMS[i], DIMS[i + 1]))
        if i < len(DIMS) - 2:
            layers.append(nn.Tanh())
    return nn.Sequential(*layers).to(device)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return False
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "samples": len(sm)}
        return False


def cpu_reference_solve(rows, threads):
    """The reference's algorithm on host cores: oracle port (torch-CPU restatement pinned bit-exact to the
    executed reference).  Returns seconds for one solve of `rows` samples."""
    from oracle import erk_oracle as eo
    torch.set_num_threads(threads)
    net = make_net("cpu")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows, DIMS[0], generator=g)
    t_span = torch.linspace(0, 1, N_STEPS_RK + 1)
    with torch.no_grad():
        t0 = time.perf_counter()
        eo.odeint(lambda t, x: net(x), x, t_span, "rk4", save_at=t_span[-1:])
        return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_solve(CPU_SAMPLE_ROWS, threads)
    times = [cpu_reference_solve(CPU_SAMPLE_ROWS, threads) for _ in range(args.steps)]
    total = sum(times)
    value = CPU_SAMPLE_ROWS * NFE_PER_SOLVE * args.steps / total
    sample = f"{CPU_SAMPLE_ROWS} of {ROWS_PER_GPU} rows per step, all {N_STEPS_RK} rk4 steps (throughput is batch-linear)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sample-NFE/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": "sample-NFE/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sample-NFE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def workload_config(world):
    return {"workload": "c2: rk4 x100 steps, MLP 32-256-256-32 tanh, forward only, save last state",
            "rows_per_gpu": ROWS_PER_GPU, "global_rows": ROWS_PER_GPU * world, "nfe_per_solve": NFE_PER_SOLVE,
            "sharding": f"batch rows x{world}, no collective (fixed-step)",
            "l2": "state + stages (>= 1.2 GB live) exceed the 126 MB L2; no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--force-generic", action="store_true", help="keep f a torch call (generic kernels only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import torchdyn_b200
    from torchdyn_b200 import _kernels, fused
    from torchdyn_b200.numerics import odeint
    if args.force_generic:
        fused.DISABLED = True
    torch.backends.cuda.matmul.allow_tf32 = False      # fp32 accuracy on the torch-f path too

    rows = args.rows
    net = make_net(dev)
    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(rows, DIMS[0], generator=g).pin_memory()
    out_host = torch.empty(rows, DIMS[0]).pin_memory()
    x_dev = x_host.to(dev)
    t_span = torch.linspace(0, 1, N_STEPS_RK + 1)
    save_at = t_span[-1:]

    def solve(x):
        with torch.no_grad():
            return odeint(net_f, x, t_span, "rk4", save_at=save_at)[1]

    model = torchdyn_b200.NeuralODE(net, solver="rk4")     # only used to standardise the vector field (DEFunc, NFE count)
    net_f = model.vf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solve(x_dev)
    # ---- device-resident timing -------------------------------------------------------------------------
    timer = _kernels.KernelTimer()
    nfe0, l0 = model.vf.nfe, _kernels.launches()
    with ClockSampler(local) as clocks:
        barrier()
        _kernels.set_timer(timer)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            sol = solve(x_dev)
        e1.record()
        barrier()
        _kernels.set_timer(None)
    ms = e0.elapsed_time(e1)
    launches = _kernels.launches() - l0
    nfe_per_solve = (model.vf.nfe - nfe0) / args.steps
    assert nfe_per_solve == NFE_PER_SOLVE, nfe_per_solve
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms)
    value = rows * world * NFE_PER_SOLVE * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host in, host result out ------------------------------------------------------
    def solve_e2e():
        xd = x_host.to(dev, non_blocking=True)
        s = solve(xd)
        out_host.copy_(s[0], non_blocking=True)

    solve_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        solve_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = rows * world * NFE_PER_SOLVE * args.steps / float(t_e)

    # ---- roofline of this library's dominant kernel ----------------------------------------------------------
    pk = peaks()
    agg = timer.summary()
    roof = None
    if agg:
        name, a = max(agg.items(), key=lambda kv: kv[1]["ms"])
        per_launch_ms = a["ms"] / a["launches"]
        if name.startswith("mlp_tc"):
            # tensor bound.  Per algorithmic FLOP the kernel issues one TF32 MMA FLOP (= 2 bf16-equivalents: TF32 dense
            # runs at half the bf16 rate) plus a bf16 correction chain over a concatenated K of twice the length
            # (a_hi*w_lo + a_lo*w_hi = 2 bf16 FLOPs): 4 bf16-equivalent FLOPs, so the ceiling is measured bf16 peak / 4.
            # (plain 3xTF32 would issue 6: its ceiling, bf16/6, is reported next to it for comparison with earlier lines)
            achieved = a["work"] / (a["ms"] * 1e-3) / 1e12
            peak = pk["bf16_sustained"] / 4.0
            traffic = None
            tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")     # from the committed ncu --set full capture
            roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "frac_of_plain_3xtf32_ceiling": achieved / (pk["bf16_sustained"] / 6.0),
                    "note": "achieved = algorithmic fp32-equivalent FLOP/s (2*sum(in*out) per sample-NFE); each is issued "
                            "as 1 TF32 MMA (hi*hi) + a bf16 MMA chain of double K (hi*lo + lo*hi) = 4 bf16-equivalent "
                            f"tensor FLOPs, so the ceiling is bf16_sustained/4 of {pk['src']}"}
        else:
            achieved = a["work"] / (a["ms"] * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": achieved / pk["hbm"], "traffic": None, "note": f"peak: {pk['src']}"}
        roof["avg_launch_ms"] = per_launch_ms
        roof["share_of_step"] = a["ms"] / ms
        roof["kernels"] = {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in agg.items()}

    # ---- CPU baseline on this box's host cores (rank 0, bounded sample) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_reference_solve(2048, threads)   # warm-up
        secs = cpu_reference_solve(CPU_SAMPLE_ROWS, threads)
        cpu = {"value": CPU_SAMPLE_ROWS * NFE_PER_SOLVE / secs, "unit": "sample-NFE/s", "cores": threads, "kind": "port",
               "sample": f"{CPU_SAMPLE_ROWS} rows x {N_STEPS_RK} rk4 steps, one solve ({secs:.1f} s)"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "sample-NFE/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "sample-NFE/s", "h2d_bytes_per_step": rows * DIMS[0] * 4,
                    "d2h_bytes_per_step": rows * DIMS[0] * 4},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks.result,
            "alg_tflops": value * FLOPS_PER_SAMPLE_NFE / 1e12,
            "path": "generic (f = torch)" if (args.force_generic or not fused.last_used()) else fused.last_used(),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
