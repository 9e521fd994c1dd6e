#!/usr/bin/env python
"""Headline benchmark of the batched ODE-integration hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--configs 1,2a,3,4 | none]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1], "c2"): rk4 fixed-step, 100 steps over t in [0, 1], MLP vector field
Linear(32,256)-Tanh-Linear(256,256)-Tanh-Linear(256,32) (random init, seed 0), synthetic state x ~ N(0,1) of
2^20 rows per GPU (weak scaling: every rank integrates its own 2^20-row shard, no data-path collective -- rk4 is
fixed-step), forward only, only the final state saved.  One "step" = one full solve = 400 vector-field
evaluations per sample.  metric = sample-NFE/s = rows * 400 / time, summed over ranks.

ONE JSON line (rank 0): value (inputs resident in HBM), e2e (pinned host buffers in, host result out, copies inside the
timed region), roofline (dominant kernel, CUDA-event timed inside the timed region), cpu_baseline (the CPU oracle port on
a bounded sub-batch), gpu_eager_baseline (the oracle's op sequence with CUDA tensors: "same box, no custom kernel"),
clocks, gpu_launches -- and `configs`: the metric's own workloads measured in the same run,
  N = 1:  c1 / c2a / c2s (= c2a at 16 384 rows: the one-launch adaptive tensor-core integrator) / c3 / c4 = dopri5 (tsit5) forward + adjoint training iterations (BASELINE.json configs[0], the
          config-2 network with dopri5 + backsolve adjoint, configs[2], configs[3]), each with ms_per_iter,
          sample_nfe_per_s, nfe, roofline, cpu_baseline and gpu_eager_baseline;
  N > 1:  the SHARDED adaptive solves whose step controller couples the ranks: config 3 strong and weak (per-stage
          all-reduce of d(mu)/dt + error-norm exchange + gradient all-reduce), the config-2 network with dopri5 +
          backsolve adjoint, and a persistent dopri5 solve whose error norm is exchanged by NVLink peer stores inside the
          running kernel; each line names its collective and the payload bytes per iteration.
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
TENSOR_KERNELS = ("mlp_tc_stage", "mlp_tc_solve", "mlp_tc_adaptive", "lin_tc", "wgrad_tc")
FFMA_KERNELS = ("mlp_solve", "mlp_solve_interp", "mlp_ffma_forward", "mlp_ffma_adjoint", "cnf_ffma_forward", "cnf_ffma_adjoint")


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
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, src="fallback (B200_PROFILING.md)")


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


# ----------------------------------------------------------------------------------------------------------------------
# the reference's algorithm off this library's kernels: CPU oracle port, and the same op sequence with CUDA tensors
# ----------------------------------------------------------------------------------------------------------------------

def oracle_c2_solve(rows, device, threads=None):
    """The reference's op sequence for c2 (oracle port: torch restatement pinned bit-exact to the executed reference) on
    `device`.  Returns seconds for one solve of `rows` samples."""
    from oracle import erk_oracle as eo
    if threads:
        torch.set_num_threads(threads)
    net = make_net(device)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows, DIMS[0], generator=g).to(device)
    t_span = torch.linspace(0, 1, N_STEPS_RK + 1)
    cuda = torch.device(device).type == "cuda"
    with torch.no_grad():
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        eo.odeint(lambda t, x: net(x), x, t_span.to(device), "rk4", save_at=t_span[-1:].to(device))
        if cuda:
            torch.cuda.synchronize()
        return time.perf_counter() - t0


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        oracle_c2_solve(CPU_SAMPLE_ROWS, "cpu", threads)
    times = [oracle_c2_solve(CPU_SAMPLE_ROWS, "cpu", threads) for _ in range(args.steps)]
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


# ----------------------------------------------------------------------------------------------------------------------
# roofline bookkeeping shared by the headline and the configs block
# ----------------------------------------------------------------------------------------------------------------------

def roofline_from(agg, step_ms, pk, traffic=None):
    """Roofline entry of the dominant kernel family in `agg` (KernelTimer.summary()) over a region of `step_ms`."""
    if not agg:
        return None
    fam = {"tensor": [k for k in agg if k in TENSOR_KERNELS], "fp32": [k for k in agg if k in FFMA_KERNELS],
           "hbm": [k for k in agg if k not in TENSOR_KERNELS and k not in FFMA_KERNELS]}
    ms_of = {b: sum(agg[k]["ms"] for k in ks) for b, ks in fam.items()}
    bound = max(ms_of, key=ms_of.get)
    ks = fam[bound]
    ms, work, launches = ms_of[bound], sum(agg[k]["work"] for k in ks), sum(agg[k]["launches"] for k in ks)
    if bound == "tensor":
        # per algorithmic FLOP the kernels issue one TF32 MMA FLOP (= 2 bf16-equivalents: TF32 dense runs at half the bf16
        # rate) plus a bf16 correction chain over a concatenated K of twice the length (a_hi*w_lo + a_lo*w_hi): 4
        # bf16-equivalent FLOPs, so the ceiling is the measured bf16 peak / 4 (plain 3xTF32 would issue 6)
        achieved, peak, unit = work / (ms * 1e-3) / 1e12, pk["bf16_sustained"] / 4.0, "TFLOP/s"
        note = ("achieved = algorithmic fp32-equivalent FLOP/s (2*sum(in*out) per forward sample-NFE, x3 for an adjoint NFE); "
                "issued as 1 TF32 MMA + a bf16 MMA chain of double K = 4 bf16-equivalent tensor FLOPs: ceiling = "
                f"bf16_sustained/4 of {pk['src']}")
    elif bound == "fp32":
        nominal = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        achieved, unit = work / (ms * 1e-3) / 1e12, "TFLOP/s"
        peak = pk.get("ffma_measured") or nominal
        note = ("narrow nets (width < 256) run as FFMA, not tensor work (north_star): ceiling = the FP32 FMA rate measured on this "
                f"GPU in this run (tdyn_probe_ffma, register-operand FFMA chains on a full grid: {peak:.1f} TFLOP/s; nominal 148 SM x "
                f"128 lanes x 2 x {pk['sm_max_mhz']:.0f} MHz = {nominal:.1f}; no driver-written peak exists for it)")
    else:
        achieved, peak, unit = work / (ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
        note = f"algorithmic bytes of the element-wise RK kernels; peak: {pk['src']}"
    return {"bound": bound, "kernel": "+".join(sorted(ks)), "achieved": achieved, "peak": peak, "unit": unit,
            "frac": achieved / peak, "traffic": traffic, "avg_launch_ms": ms / max(launches, 1),
            "share_of_step": ms / step_ms, "note": note,
            "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in agg.items()}}


# ----------------------------------------------------------------------------------------------------------------------
# configs block, one GPU: the metric's own workloads (forward + adjoint training iterations)
# ----------------------------------------------------------------------------------------------------------------------
CONFIG_IDS = {"1": 1, "2a": 2, "2s": 2, "3": 3, "4": 4, "5": 5}
# "2s": config 2a at 1/16 of the batch (16 384 rows) -- the size class of the ONE-LAUNCH adaptive tensor-core integrator
# (tdyn_mlp_tc_solve_adaptive; above 32 768 rows the host-driven loop is faster and is what "2a" measures)
CONFIG_SCALE = {"2s": 1.0 / 16}
CPU_SCALE = {1: 1.0, 2: 1 / 32, 3: 1 / 32, 4: 1 / 8, 5: 1 / 32}       # sub-batch of the CPU leg (throughput is batch-linear)


def measure_config(key, dev, pk, iters=5, cpu=True, eager=True):
    import torchdyn_b200
    from benchmarks import run_configs as rc
    from torchdyn_b200 import _kernels, fused
    cfg, scale = CONFIG_IDS[key], CONFIG_SCALE.get(key, 1.0)
    net, X, t_span, kw, loss_fn, name = rc.build(cfg, dev, scale)
    net = net.to(dev)
    model = torchdyn_b200.NeuralODE(net, **kw)
    x = X.to(dev)
    x_host = X.pin_memory()
    B = x.shape[0]

    def it(xin):
        for p in net.parameters():
            p.grad = None
        xx = xin.clone().requires_grad_(True)
        _, sol = model(xx, t_span)
        loss = loss_fn(sol)
        loss.backward()
        return loss

    for _ in range(3):
        it(x)
    torch.cuda.synchronize()
    n0, l0 = model.vf.nfe, _kernels.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        it(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nfe = (model.vf.nfe - n0) / iters
    launches = (_kernels.launches() - l0) / iters
    path = fused.last_used()
    # end to end: pinned host batch in, loss value out, every iteration
    w0 = time.perf_counter()
    for _ in range(iters):
        float(it(x_host.to(dev, non_blocking=True)).item())
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - w0) / iters
    # kernel breakdown of ONE iteration (events around every launch of this library)
    timer = _kernels.KernelTimer()
    _kernels.set_timer(timer)
    e0.record()
    it(x)
    e1.record()
    _kernels.set_timer(None)
    agg = timer.summary()
    out = {"workload": name, "batch": B, "nfe_per_iter": nfe, "ms_per_iter": ms, "sample_nfe_per_s": B * nfe / (ms * 1e-3),
           "e2e": {"ms_per_iter": e2e_ms, "value": B * nfe / (e2e_ms * 1e-3), "unit": "sample-NFE/s",
                   "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4},
           "gpu_launches_per_iter": launches, "path": path, "roofline": roofline_from(agg, e0.elapsed_time(e1), pk)}
    del model
    if eager:
        out["gpu_eager_baseline"] = eager_config(cfg, dev, B, nfe, out["sample_nfe_per_s"], scale)
    if cpu:
        try:
            r = rc.run_cpu(cfg, 1, min(CPU_SCALE[cfg], scale))
            out["cpu_baseline"] = {"value": r["sample_nfe_per_s"], "unit": "sample-NFE/s", "cores": r["cores"], "kind": "port",
                                   "sample": f"{r['batch']} of {B} rows, one training iteration ({r['ms_per_iter'] / 1e3:.1f} s, "
                                             f"{r['nfe_per_iter']:.0f} NFE)"}
        except Exception as e:      # the CPU leg must never take the GPU numbers down with it
            out["cpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"}
    torch.cuda.empty_cache()
    return out


def eager_config(cfg, dev, B, nfe_b200, value_b200, scale=1.0):
    """The oracle's op sequence (= the reference's: ~78 ATen launches per dopri5 step, autograd VJPs) with CUDA tensors."""
    from benchmarks import run_configs as rc
    from oracle import erk_oracle as eo
    try:
        net, X, t_span, kw, loss_fn, _ = rc.build(cfg, dev, scale)
        net = net.to(dev)
        x = X.to(dev)
        ts = t_span.to(dev)
        if getattr(net, "noise_dist", None) is not None:      # CNF: NeuralODE samples the Hutchinson noise once per forward
            net.noise = net.noise_dist.sample((x.shape[0],))
        model = eo.OracleNeuralODE(net, **kw)

        def it():
            for p in net.parameters():
                p.grad = None
            xx = x.clone().requires_grad_(True)
            _, sol = model(xx, ts)
            loss_fn(sol).backward()

        it()
        torch.cuda.synchronize()
        n0 = model.nfe
        t0 = time.perf_counter()
        for _ in range(2):
            it()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / 2
        nfe = (model.nfe - n0) / 2
        v = B * nfe / (ms * 1e-3)
        return {"value": v, "unit": "sample-NFE/s", "ms_per_iter": ms, "nfe_per_iter": nfe, "speedup": value_b200 / v,
                "what": "oracle port (reference op sequence, torch eager ATen kernels, fp32 no-TF32) with tensors on this GPU"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}
    finally:
        torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------------------------------------
# configs block, N > 1 GPUs: sharded adaptive solves (the rank-coupling collectives of SURVEY section 8e)
# ----------------------------------------------------------------------------------------------------------------------

def sharded_configs(dev, rank, world, iters=3):
    import torch.distributed as dist
    import torchdyn_b200
    from torchdyn_b200 import distributed as tdist, fused
    from torchdyn_b200.numerics import odeint
    out = {}
    t_span = torch.linspace(0, 1, 2)

    def timed(step):
        for _ in range(2):
            step()
        dist.barrier()
        torch.cuda.synchronize()
        tdist.traffic(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            nfe = step()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        tr = tdist.traffic(reset=True)
        return float(ms), nfe, {k: v / iters for k, v in tr.items()}

    def training(net, shard, B_global, kw, label, collective):
        def step():
            for p in net.parameters():
                p.grad = None
            tdist.enable()
            model = torchdyn_b200.NeuralODE(net, **kw)
            _, sol = model(shard, t_span)
            (sol[-1].pow(2).sum() / (B_global * shard.shape[1])).backward()
            tdist.allreduce_gradients(net.parameters())
            tdist.disable()
            return float(model.vf.nfe)
        ms, nfe, tr = timed(step)
        return {"workload": label, "global_batch": B_global, "rows_per_gpu": shard.shape[0], "nfe_per_iter": nfe,
                "ms_per_iter": ms, "sample_nfe_per_s": B_global * nfe / (ms * 1e-3), "sharding": collective,
                "nvlink_peer_exchanges_per_iter": tr["peer_calls"], "nvlink_peer_payload_bytes_per_iter_per_rank": tr["peer_bytes"],
                "nccl_collectives_per_iter": tr["calls"], "nccl_payload_bytes_per_iter_per_rank": tr["bytes"],
                "path": fused.last_used()}

    # ---- config 3 (tsit5 + interpolated adjoint, 128-1024-128): strong (262 144 rows in total) and weak (per GPU)
    torch.manual_seed(0)
    net3 = nn.Sequential(nn.Linear(128, 1024), nn.Tanh(), nn.Linear(1024, 128)).to(dev)
    kw3 = dict(solver="tsit5", sensitivity="interpolated_adjoint")
    coll3 = (f"batch rows x{world}; per attempted step an all-reduce(SUM) of 4 fp64 (global RMS error norm: sum of squares + "
             "counts) by tdyn_peer_allreduce = ONE kernel per GPU that all-gathers by direct NVLink peer stores into every "
             "rank's exchange buffer, releases per-rank flags and adds in rank order (deterministic, no NCCL call); per "
             "adjoint stage an NCCL all-reduce(SUM) of d(mu)/dt (263 296 fp32: mu sits inside the interpolated adjoint's "
             "norm; the NVSwitch reduces it faster than peer stores do); one gradient all-reduce(SUM) per iteration (NCCL)")
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    rows = 262144 // world
    out["c3_strong"] = training(net3, torch.randn(rows, 128, generator=g, device=dev), 262144, kw3,
                                "c3 tsit5 + interpolated adjoint, MLP 128-1024-128, global batch 262144 (strong scaling)", coll3)
    out["c3_weak"] = training(net3, torch.randn(262144, 128, generator=g, device=dev), 262144 * world, kw3,
                              "c3 tsit5 + interpolated adjoint, MLP 128-1024-128, 262144 rows per GPU (weak scaling)", coll3)
    del net3
    torch.cuda.empty_cache()
    # ---- the config-2 network with dopri5 + backsolve adjoint (the metric's "dopri5 MLP ODE fwd+adjoint"), weak
    torch.manual_seed(0)
    net2 = make_net(dev)
    kw2 = dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint", atol_adjoint=1e-4, rtol_adjoint=1e-4)
    coll2 = (f"batch rows x{world}; per attempted step (forward and adjoint) an all-reduce(SUM) of 4 fp64 (global RMS error "
             "norm / seminorm) by tdyn_peer_allreduce (in-kernel NVLink peer stores + flags, fixed-order sum), one gradient "
             "all-reduce(SUM) of 82 464 fp32 per iteration through NCCL")
    out["c2a_weak"] = training(net2, torch.randn(262144, 32, generator=g, device=dev), 262144 * world, kw2,
                               "c2a dopri5 1e-4 + backsolve adjoint, MLP 32-256-256-32, 262144 rows per GPU (weak scaling)", coll2)
    del net2
    # ---- persistent dopri5 solve, error norm exchanged INSIDE the running kernel by NVLink peer stores (no NCCL call)
    torch.manual_seed(0)
    netp = nn.Sequential(nn.Linear(8, 64), nn.Tanh(), nn.Linear(64, 8)).to(dev)
    with torch.no_grad():
        for p in netp.parameters():
            p.mul_(2.0)
    xp = torch.randn(262144, 8, generator=g, device=dev) * 1.5
    fld = fused.MLPField(netp)
    om = sys.modules["torchdyn_b200.numerics.odeint"]
    tsp = torch.tensor([0.0, 1.0, 2.0])
    attempts = [0]

    def solve():
        tdist.enable()
        om.TRACE_SINK = tr = []
        n0 = fld.nfe
        with torch.no_grad():
            odeint(fld, xp, tsp, "dopri5", atol=1e-5, rtol=1e-5)
        om.TRACE_SINK = None
        tdist.disable()
        attempts[0] = len(tr[0]["t"]) if tr else 0
        return float(fld.nfe - n0)
    ms, nfe, _ = timed(solve)
    path = fused.last_used() or ""
    out["dopri5_persistent_weak"] = {
        "workload": "dopri5 1e-5 forward, MLP 8-64-8 (gain 2), 262144 rows per GPU, whole solve in ONE persistent kernel per GPU",
        "global_batch": 262144 * world, "rows_per_gpu": 262144, "nfe_per_iter": nfe, "ms_per_iter": ms,
        "sample_nfe_per_s": 262144 * world * nfe / (ms * 1e-3), "attempted_steps": attempts[0],
        "sharding": (f"batch rows x{world}; per attempted step every GPU stores its fp64 partial sums into every peer's exchange "
                     "buffer over NVLink from inside the running kernel (st.release.sys / ld.acquire.sys, fixed-order sum): "
                     "all-gather + sum fused into the integrator, no NCCL call, no host round trip"),
        "nvlink_payload_bytes_per_step_per_rank": 8 * 4 * (world - 1), "path": path}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help=argparse.SUPPRESS)
    ap.add_argument("--configs", default="1,2a,2s,3,4,5", help="configs block: comma-separated of 1,2a,2s,3,4,5 or 'none'")
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-eager-baseline", action="store_true", help=argparse.SUPPRESS)
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
    torch.backends.cudnn.allow_tf32 = False

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
    path = "generic (f = torch)" if (args.force_generic or not fused.last_used()) else fused.last_used()

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
    try:
        # FP32 FMA rate measured live (register-operand FFMA chains on a full grid): the ceiling of the FFMA kernels
        pk["ffma_measured"] = _kernels.get(torch.empty(0, device=dev)).probe_ffma_tflops()
    except Exception:
        pk["ffma_measured"] = None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k2_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")     # from the committed ncu --set full capture
    roof = roofline_from(timer.summary(), ms, pk, traffic)
    if roof and roof["bound"] == "tensor":
        roof["frac_of_plain_3xtf32_ceiling"] = roof["achieved"] / (pk["bf16_sustained"] / 6.0)
    del sol, x_dev
    torch.cuda.empty_cache()

    # ---- the same solve off this library's kernels: CPU oracle port, and the oracle's op sequence on this GPU ------
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        oracle_c2_solve(2048, "cpu", threads)   # warm-up
        secs = oracle_c2_solve(CPU_SAMPLE_ROWS, "cpu", threads)
        cpu = {"value": CPU_SAMPLE_ROWS * NFE_PER_SOLVE / secs, "unit": "sample-NFE/s", "cores": threads, "kind": "port",
               "sample": f"{CPU_SAMPLE_ROWS} rows x {N_STEPS_RK} rk4 steps, one solve ({secs:.1f} s)"}
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        try:
            oracle_c2_solve(4096, dev)
            secs = oracle_c2_solve(rows, dev)
            ev = rows * NFE_PER_SOLVE / secs
            eager = {"value": ev, "unit": "sample-NFE/s", "ms_per_step": 1e3 * secs, "speedup": value / ev, "rows": rows,
                     "what": "oracle port (the reference's op sequence: 400 torch fp32 MLP calls + ~2 400 element-wise ATen "
                             "launches, no TF32) with the state on this GPU: same box, no custom kernel (BASELINE.md section 3)"}
        except Exception as e:
            eager = {"unavailable": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    # ---- configs block -----------------------------------------------------------------------------------------------
    configs = {}
    keys = [k for k in args.configs.split(",") if k and k != "none"]
    if world == 1:
        for k in keys:
            try:
                configs["c" + k] = measure_config(k, dev, pk, cpu=not args.no_cpu_baseline, eager=not args.no_eager_baseline)
            except Exception as e:
                configs["c" + k] = {"error": f"{type(e).__name__}: {e}"}
    elif keys:
        try:
            configs = sharded_configs(dev, rank, world)
        except Exception as e:
            configs = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "sample-NFE/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "sample-NFE/s", "h2d_bytes_per_step": rows * DIMS[0] * 4,
                    "d2h_bytes_per_step": rows * DIMS[0] * 4},
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
            "clocks": clocks.result, "alg_tflops": value * FLOPS_PER_SAMPLE_NFE / 1e12, "path": path, "configs": configs,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
