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

_state = {"enabled": False, "group": None, "symm": None, "epoch": 0, "symm_failed": False, "replicated_from": None}
_traffic = {"calls": 0, "bytes": 0, "peer_calls": 0, "peer_bytes": 0}   # collectives issued by this module on this rank
                                                                          # (NCCL calls / in-kernel NVLink exchanges)
PEER_ALLREDUCE = True     # all-reduce small vectors inside one kernel per GPU over peer-mapped memory (False: NCCL)
PEER_MAX_BYTES = 32768    # larger payloads go to NCCL: measured on 4 / 8 B200 (benchmarks/multi_gpu_peer_allreduce.py), 4 doubles
                          # take 15 / 25 us in the peer kernel vs 19 / 34 us in NCCL, but 1 MB takes 37 / 59 us vs 20 / 29 us
                          # (NCCL reduces inside the NVSwitch; the peer kernel's flag handshakes cost ~15 us per phase)


def traffic(reset: bool = False):
    """(number of collectives, payload bytes) this rank issued through this module since the last reset."""
    out = dict(_traffic)
    if reset:
        for k in _traffic:
            _traffic[k] = 0
    return out


def _count(t: torch.Tensor):
    _traffic["calls"] += 1
    _traffic["bytes"] += t.numel() * t.element_size()


def enable(group=None):
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["enabled"], _state["group"] = True, group


def disable():
    _state["enabled"], _state["group"] = False, None


def peer_comm(device):
    """tdyn_comm_t for the in-kernel NVLink exchange of the persistent integrator, or None when peer-mapped
    (symmetric) memory is unavailable -- callers then use the host-driven loop with the NCCL all-reduce.
    Collective on first use (symmetric-memory rendezvous); every later call must be made by all ranks in the same
    order (it advances the launch epoch)."""
    if not _state["enabled"] or _state["symm_failed"] or device.type != "cuda":
        return None
    import torch.distributed as dist
    from . import _lib
    if _state["symm"] is None:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            lib = _lib.load()
            nbytes = int(lib.tdyn_comm_buffer_bytes())
            group = _state["group"] if _state["group"] is not None else dist.group.WORLD
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group)
            torch.cuda.synchronize(device)
            dist.barrier(group)
            if hdl.world_size > 8:
                raise RuntimeError("more than 8 ranks")
            _state["symm"] = (buf, hdl, [int(p) for p in hdl.buffer_ptrs])
        except Exception as e:       # no peer access / API unavailable: stay on the NCCL path
            import warnings
            warnings.warn(f"torchdyn_b200: symmetric memory unavailable ({e}); sharded solves use the NCCL all-reduce path")
            _state["symm_failed"] = True
            return None
    buf, hdl, ptrs = _state["symm"]
    _state["epoch"] += 1
    c = _lib.Comm()
    c.world, c.rank, c.epoch = hdl.world_size, hdl.rank, _state["epoch"]
    for r, p in enumerate(ptrs):
        c.peer[r] = p
    return c


def is_enabled() -> bool:
    return _state["enabled"]


def world_size() -> int:
    if not _state["enabled"]:
        return 1
    import torch.distributed as dist
    return dist.get_world_size(_state["group"])


def rank() -> int:
    if not _state["enabled"]:
        return 0
    import torch.distributed as dist
    return dist.get_rank(_state["group"])


class replicated_tail:
    """``with replicated_tail(n):`` -- inside, the flat state integrated by ``odeint`` is ``[sharded (n elements of
    this rank), replicated]``: the tail is identical on every rank (the interpolated adjoint's mu block once its stage
    derivatives are all-reduced, reference sensitivity.py:124,152 puts mu INSIDE the error norm), so only rank 0
    contributes it to the global sums of squares and element counts."""

    def __init__(self, n_sharded: int):
        self.n = int(n_sharded)

    def __enter__(self):
        self.prev = _state["replicated_from"]
        if _state["enabled"]:
            _state["replicated_from"] = self.n
        return self

    def __exit__(self, *exc):
        _state["replicated_from"] = self.prev
        return False


def local_norm_count(count: int) -> int:
    """Of the first ``count`` elements of this rank's flat state, how many this rank contributes to a global norm."""
    rf = _state["replicated_from"]
    if not _state["enabled"] or rf is None or rank() == 0:
        return int(count)
    return min(int(count), rf)


def _peer_allreduce_(t: torch.Tensor) -> bool:
    """In-place sum over ranks by ``tdyn_peer_allreduce`` (one kernel per GPU: all-gather by NVLink peer stores, flags,
    fixed-order sum).  False when unavailable (no peer-mapped memory, more than 8 ranks, unsupported tensor)."""
    if not PEER_ALLREDUCE or _state.get("ar_failed") or not (t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.float64)):
        return False
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    nbytes = t.numel() * t.element_size()
    if nbytes > PEER_MAX_BYTES:
        return False
    ar = _state.get("ar")
    if ar is None or nbytes > ar["slot_bytes"]:
        try:
            import torch.distributed._symmetric_memory as symm_mem
            lib = _lib.load()
            group = _state["group"] if _state["group"] is not None else dist.group.WORLD
            slot = max(2 << 20, (nbytes + (1 << 20) - 1) >> 20 << 20)          # collective: every rank computes the same size
            buf = symm_mem.empty(int(lib.tdyn_peer_allreduce_buffer_bytes(slot)), dtype=torch.uint8, device=t.device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group)
            torch.cuda.synchronize(t.device)
            dist.barrier(group)
            if hdl.world_size > 8:
                raise RuntimeError("more than 8 ranks")
            ar = _state["ar"] = {"buf": buf, "hdl": hdl, "ptrs": [int(p) for p in hdl.buffer_ptrs], "slot_bytes": slot, "seq": 0}
        except Exception as e:
            import warnings
            warnings.warn(f"torchdyn_b200: peer-memory all-reduce unavailable ({e}); using NCCL")
            _state["ar_failed"] = True
            return False
    ar["seq"] += 1
    c = _lib.Comm()
    c.world, c.rank, c.epoch = ar["hdl"].world_size, ar["hdl"].rank, ar["seq"]
    for r, p in enumerate(ar["ptrs"]):
        c.peer[r] = p
    with torch.cuda.device(t.device):
        rc = _lib.load().tdyn_peer_allreduce(C.byref(c), t.data_ptr(), t.numel(), int(t.dtype == torch.float64), ar["slot_bytes"],
                                             torch.cuda.current_stream(t.device).cuda_stream)
    _lib.check(rc, "tdyn_peer_allreduce")
    _traffic["peer_calls"] += 1
    _traffic["peer_bytes"] += nbytes
    return True


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op when sharding is off)."""
    if _state["enabled"]:
        if _peer_allreduce_(t):
            return t
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_state["group"])
        _count(t)
    return t


def reduce_sums(kern, n_sums: int, local_count: int) -> Tuple[List[float], int]:
    """Read the first ``n_sums`` fp64 reduction results of ``kern`` (summed over ranks when sharding is on)
    and the matching global element count.  This is the step's single host sync."""
    if not _state["enabled"]:
        return kern.read_reduction(n_sums), int(local_count)
    import torch.distributed as dist
    kern.red[3].fill_(float(local_count))
    if not _peer_allreduce_(kern.red):
        dist.all_reduce(kern.red, op=dist.ReduceOp.SUM, group=_state["group"])
        _count(kern.red)
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
    _count(flat)
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
