#!/usr/bin/env python
"""Sharded adaptive FORWARD solve in the persistent kernel: the per-step global error norm is exchanged between the
GPUs by direct NVLink peer stores inside the kernel (no NCCL call, no host round trip per step).  Rank 0 also runs the
unsharded problem on its GPU and checks that the step sequence is the same.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/multi_gpu_persistent.py
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
    import torchdyn_b200  # noqa: F401
    from torchdyn_b200 import distributed as tdist, fused
    from torchdyn_b200.numerics import odeint
    om = sys.modules["torchdyn_b200.numerics.odeint"]

    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(8, 64), nn.Tanh(), nn.Linear(64, 8)).to(dev)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(2.0)
    g = torch.Generator().manual_seed(1)
    B = 262144 + 5
    x_full = (torch.randn(B, 8, generator=g) * 1.5).to(dev)
    t_span = torch.tensor([0.0, 1.0, 2.0])
    fld = fused.MLPField(net)

    def run(x, sharded, persistent=True):
        (tdist.enable if sharded else tdist.disable)()
        fused.PERSISTENT = persistent
        om.TRACE_SINK = tr = []
        torch.cuda.synchronize()
        if sharded:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.no_grad():
            t_eval, sol = odeint(fld, x, t_span, "dopri5", atol=1e-5, rtol=1e-5)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        om.TRACE_SINK = None
        path = fused.last_used()
        fused.PERSISTENT = True
        tdist.disable()
        return tr[0], sol, dt, path

    shard = tdist.shard(x_full, rank, world).contiguous()
    run(shard, True)
    tr_s, sol_s, dt_s, path_s = run(shard, True)
    _, _, dt_nccl, path_n = run(shard, True, persistent=False)          # host loop + NCCL all-reduce per step, for comparison
    if rank == 0:
        run(x_full, False)
        tr_f, sol_f, dt_f, _ = run(x_full, False)
        same = tr_s["accept"] == tr_f["accept"] and len(tr_s["t"]) == len(tr_f["t"])
        t_dev = max(abs(u - v) / max(abs(v), 1e-3) for u, v in zip(tr_s["t"], tr_f["t"]))
        mine = tdist.shard(sol_f[-1], 0, world)
        s_dev = float((sol_s[-1] - mine).abs().max() / mine.abs().max())
        print(json.dumps({"world": world, "rows_total": B, "path_sharded": path_s, "same_step_sequence": same,
                          "attempts": len(tr_s["t"]), "max_rel_dev_t": t_dev, "max_rel_dev_state": s_dev,
                          "ms_sharded_peer_store_kernel": 1e3 * dt_s, "ms_sharded_host_loop_nccl": 1e3 * dt_nccl,
                          "ms_single_gpu_full_batch": 1e3 * dt_f}), flush=True)
        assert "peer stores" in (path_s or ""), path_s
        assert same and t_dev < 1e-5 and s_dev < 1e-5
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
