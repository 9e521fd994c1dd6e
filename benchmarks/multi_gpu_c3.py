#!/usr/bin/env python
"""BASELINE config 3 sharded over N GPUs (torchrun): tsit5 + interpolated adjoint, MLP 128-1024-128, global batch
262 144, one training step = forward + backward + gradient all-reduce.  The vector field, its VJP and the parameter
gradient run on the layer-wise tensor-core kernels; per attempted step the ranks all-reduce 4 doubles (error norm) and,
in the backward, the d(mu)/dt block of every stage (the mu block is inside the interpolated adjoint's error norm).
Rank 0 also runs the unsharded step and compares NFE and gradients.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/multi_gpu_c3.py
"""
import json
import os
import sys

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
    from torchdyn_b200 import distributed as tdist, fused

    weak = "--weak" in sys.argv                 # 262 144 rows PER GPU instead of in total
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(128, 1024), nn.Tanh(), nn.Linear(1024, 128)).to(dev)
    B = 262144 * (world if weak else 1)
    if weak:
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        x_full = None
    else:
        g = torch.Generator().manual_seed(1)
        x_full = torch.randn(B, 128, generator=g).to(dev)
    t_span = torch.linspace(0, 1, 2)

    def step(x, sharded):
        for p in net.parameters():
            p.grad = None
        (tdist.enable if sharded else tdist.disable)()
        model = torchdyn_b200.NeuralODE(net, solver="tsit5", sensitivity="interpolated_adjoint")
        _, sol = model(x, t_span)
        (sol[-1].pow(2).sum() / (B * 128)).backward()
        if sharded:
            tdist.allreduce_gradients(net.parameters())
        nfe = float(model.vf.nfe)
        tdist.disable()
        return nfe

    def timed(x, sharded, iters=3):
        step(x, sharded)
        step(x, sharded)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            nfe = step(x, sharded)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        return ms, nfe

    shard = (torch.randn(262144, 128, generator=g, device=dev) if weak else tdist.shard(x_full, rank, world).contiguous())
    ms, nfe_s = timed(shard, True)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    g_s = [p.grad.clone() for p in net.parameters()]
    path = fused.last_used()
    if rank == 0 and weak:
        print(json.dumps({"config": "c3 tsit5 + interpolated adjoint, MLP 128-1024-128, 262144 rows per GPU (weak scaling)",
                          "world": world, "global_batch": B, "ms_per_iter_sharded_max_over_ranks": float(ms),
                          "sample_nfe_per_s": B * nfe_s / (float(ms) * 1e-3), "nfe_sharded": nfe_s, "path": path}), flush=True)
    elif rank == 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        step(x_full, False)
        e0.record()
        nfe_f = step(x_full, False)
        e1.record()
        torch.cuda.synchronize()
        g_f = [p.grad.clone() for p in net.parameters()]
        g_dev = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(g_s, g_f))
        print(json.dumps({"config": "c3 tsit5 + interpolated adjoint, MLP 128-1024-128, global batch 262144 (strong scaling)",
                          "world": world, "ms_per_iter_sharded_max_over_ranks": float(ms), "ms_per_iter_single_gpu": e0.elapsed_time(e1),
                          "sample_nfe_per_s": B * nfe_s / (float(ms) * 1e-3), "nfe_sharded": nfe_s, "nfe_single": nfe_f,
                          "max_rel_dev_grad_vs_single_gpu": g_dev, "path": path}), flush=True)
        assert nfe_s == nfe_f and g_dev < 2e-3
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
