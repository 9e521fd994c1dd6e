#!/usr/bin/env python
"""In-kernel all-reduce over peer-mapped memory (tdyn_peer_allreduce) against NCCL: values (fixed-order sum of the ranks'
vectors, bit for bit) and time per call, for the two payloads of the sharded solves (the d(mu)/dt block of config 3 and the
4-double error-norm sums).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 benchmarks/multi_gpu_peer_allreduce.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from torchdyn_b200 import distributed as tdist
    tdist.enable()
    tdist.PEER_MAX_BYTES = 1 << 30          # measure the peer kernels at every size (the library sends > 32 KB to NCCL)
    out = {"world": world}
    for label, n, dtype in (("dmu_263296_f32", 263296, torch.float32), ("sums_4_f64", 4, torch.float64), ("odd_100003_f32", 100003, torch.float32)):
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        ok = True
        for it in range(20):
            v = torch.randn(n, generator=g, device=dev, dtype=dtype)
            gathered = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(gathered, v)
            expect = gathered[0].clone()
            for q in range(1, world):
                expect += gathered[q]
            w = v.clone()
            assert tdist._peer_allreduce_(w), "peer all-reduce unavailable"
            ok = ok and torch.equal(w, expect)
        times = {}
        for name, fn in (("peer_kernel", lambda t: tdist._peer_allreduce_(t)), ("nccl", lambda t: dist.all_reduce(t))):
            v = torch.randn(n, device=dev, dtype=dtype)
            for _ in range(10):
                fn(v)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn(v)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[name] = float(t) * 1e3
        okt = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        out[label] = {"bit_identical_to_rank_ordered_sum": bool(int(okt)), "us_per_call_peer_kernel": times["peer_kernel"],
                      "us_per_call_nccl": times["nccl"]}
    if rank == 0:
        print(json.dumps(out), flush=True)
    assert all(v["bit_identical_to_rank_ordered_sum"] for k, v in out.items() if isinstance(v, dict))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
