"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no distributed code; its adaptive controller is ONE global RMS norm over the whole batch
(torchdyn/numerics/utils.py:33-34, numerics/odeint.py:371-373).  Sharding the batch keeps every RK stage
rank-local; the only coupling is that norm.  Each rank reduces its shard to a fp64 sum of squares on the
device, the sums (and the element counts, so unequal shards work) are added across ranks with a single
all-reduce(SUM) of 4 doubles per attempted step, and every rank then takes the identical accept / dt decision.
(`north_star` says all-reduce(max); a max of per-shard RMS values would NOT reproduce the reference's
controller -- the sum of squares does.)  Fixed-step solvers need no collective at all.

    torchdyn_b200.distributed.enable(group=None)    # after dist.init_process_group(...)
    ... odeint / NeuralODE on this rank's shard ...
    torchdyn_b200.distributed.allreduce_gradients(model)   # sum of per-shard adjoint gradients
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch

_state = {"enabled": False, "group": None}


def enable(group=None):
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["enabled"], _state["group"] = True, group


def disable():
    _state["enabled"], _state["group"] = False, None


def is_enabled() -> bool:
    return _state["enabled"]


def world_size() -> int:
    if not _state["enabled"]:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(_state["group"])


def reduce_sums(kern, n_sums: int, local_count: int) -> Tuple[List[float], int]:
    """Read the first ``n_sums`` fp64 reduction results of ``kern`` (summed over ranks when sharding is on)
    and the matching global element count.  This is the step's single host sync."""
    if not _state["enabled"]:
        return kern.read_reduction(n_sums), int(local_count)
    import torch.distributed as dist
    kern.red[3].fill_(float(local_count))
    dist.all_reduce(kern.red, op=dist.ReduceOp.SUM, group=_state["group"])
    vals = kern.read_reduction(4)
    return vals[:n_sums], int(round(vals[3]))


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = False):
    """Sum (or average) ``.grad`` over ranks in ONE flat all-reduce (the per-shard adjoint gradients mu are
    partial sums over the batch, sensitivity.py:95-98 in the reference)."""
    if not _state["enabled"]:
        return
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_state["group"])
    if average:
        flat /= world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def shard(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's contiguous slice of the batch dimension (unequal remainders go to the first ranks)."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank(_state["group"])
    if world is None:
        world = world_size()
    n = x.shape[0]
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return x[start:start + base + (1 if rank < rem else 0)]
