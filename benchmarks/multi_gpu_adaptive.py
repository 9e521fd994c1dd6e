#!/usr/bin/env python
"""Batch-sharded ADAPTIVE solve + adjoint on N GPUs (torchrun): every rank integrates its shard, the per-step
all-reduce of the squared-error sums makes all ranks take the single global controller's steps.  Rank 0 also runs the
unsharded problem and checks that the step sequence and the gradients agree.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/multi_gpu_adaptive.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torchdyn_b200
    from torchdyn_b200 import distributed as tdist
    om = sys.modules["torchdyn_b200.numerics.odeint"]

    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(8, 64), nn.Tanh(), nn.Linear(64, 8)).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(2.0)
    g = torch.Generator().manual_seed(1)
    B = 65536 + 3                                           # unequal shards
    x_full = (torch.randn(B, 8, generator=g) * 1.5).to(dev)
    t_span = torch.tensor([0.0, 2.0])
    kw = dict(solver="dopri5", atol=1e-5, rtol=1e-5, sensitivity="adjoint", atol_adjoint=1e-5, rtol_adjoint=1e-5)

    from torchdyn_b200 import fused

    def run(x, sharded, persistent=True):
        for p in net.parameters():
            p.grad = None
        fused.PERSISTENT = persistent
        (tdist.enable if sharded else tdist.disable)()
        model = torchdyn_b200.NeuralODE(net, **kw)
        om.TRACE_SINK = tr = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, sol = model(x.clone().requires_grad_(True), t_span)
        (sol[-1].pow(2).sum() / B).backward()
        if sharded:
            tdist.allreduce_gradients(net.parameters())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        om.TRACE_SINK = None
        tdist.disable()
        fused.PERSISTENT = True
        return tr, [p.grad.clone() for p in net.parameters()], dt, model.vf.nfe

    shard = tdist.shard(x_full, rank, world).contiguous()
    run(shard, True)                                        # warm-up (persistent kernels, NVLink peer exchange)
    tr_s, g_s, dt_s, nfe = run(shard, True)
    path = fused.last_used()
    run(shard, True, persistent=False)                      # host-driven loops, NCCL all-reduce per attempted step
    tr_h, g_h, dt_h, _ = run(shard, True, persistent=False)
    if rank == 0:
        run(x_full, False)
        tr_f, g_f, dt_f, _ = run(x_full, False)
        same_steps = all(a["accept"] == b["accept"] and len(a["t"]) == len(b["t"]) for a, b in zip(tr_s, tr_f))
        t_dev = max(abs(u - v) / max(abs(v), 1e-3) for a, b in zip(tr_s, tr_f) for u, v in zip(a["t"], b["t"]))
        g_dev = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(g_s, g_f))
        print(json.dumps({"world": world, "rows_total": B, "nfe": nfe, "same_step_sequence": same_steps,
                          "max_rel_dev_t": t_dev, "max_rel_dev_grad": g_dev, "path_sharded": path,
                          "ms_sharded_persistent_peer_exchange": 1e3 * dt_s, "ms_sharded_host_loop_nccl": 1e3 * dt_h,
                          "same_steps_host_loop": all(a["accept"] == b["accept"] for a, b in zip(tr_h, tr_f)),
                          "ms_single_gpu_full_batch": 1e3 * dt_f, "attempts_per_call": [len(a["t"]) for a in tr_s]}), flush=True)
        assert same_steps and t_dev < 2e-3 and g_dev < 1e-4   # dt inherits the noise of error ratios (different f kernels / reduction shapes)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
