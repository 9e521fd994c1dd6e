"""World-size-2 gloo test of the batch-sharded path on CPU (host logic only; kernels replaced by the test double).

Each rank integrates half of the batch with torchdyn_b200.distributed enabled: the per-shard sums of squared
scaled errors (and element counts) are all-reduced once per attempted step, so both ranks must take exactly the
step sequence of the single global controller that an unsharded run over the whole batch takes."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _solve(x, t_span, solver, traces):
    import torchdyn_b200  # noqa: F401
    from torchdyn_b200.numerics import odeint
    om = sys.modules["torchdyn_b200.numerics.odeint"]
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(4, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(3.0)
    om.TRACE_SINK = traces
    try:
        with torch.no_grad():
            return odeint(lambda t, x: net(x), x, t_span, solver, atol=1e-4, rtol=1e-4)
    finally:
        om.TRACE_SINK = None


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from torchdyn_b200 import _kernels, distributed
    from cpu_kernel_double import CpuKernels
    k = CpuKernels()
    _kernels.set_test_backend(lambda x: k)
    distributed.enable()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(101, 4, generator=g) * 2          # odd size: unequal shards (51 + 50)
    shard = distributed.shard(x, rank, world).contiguous()
    traces = []
    t_eval, sol = _solve(shard, torch.tensor([0.0, 3.0]), "dopri5", traces)
    torch.save(dict(t=traces[0]["t"], dt=traces[0]["dt"], ratio=traces[0]["ratio"], sol=sol[-1], n=shard.shape[0]),
               os.path.join(out_dir, f"rank{rank}.pt"))
    # gradient all-reduce helper: sum of per-rank grads
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), float(rank + 1))
    distributed.allreduce_gradients([p])
    assert torch.equal(p.grad, torch.full((3,), 3.0))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_adaptive_solve_takes_the_global_controller_steps(tmp_path):
    sys.path.insert(0, HERE)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    assert r0["n"] == 51 and r1["n"] == 50
    # both ranks took identical decisions
    assert r0["t"] == r1["t"] and r0["dt"] == r1["dt"] and r0["ratio"] == r1["ratio"]

    # unsharded run over the whole batch in this process
    from torchdyn_b200 import _kernels
    from cpu_kernel_double import CpuKernels
    torch.set_num_threads(1)
    k = CpuKernels()
    _kernels.set_test_backend(lambda x: k)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(101, 4, generator=g) * 2
        traces = []
        _, sol = _solve(x, torch.tensor([0.0, 3.0]), "dopri5", traces)
    finally:
        _kernels.set_test_backend(None)
    full = traces[0]
    assert len(full["t"]) == len(r0["t"]) and full["accept"] == [r <= 1 for r in r0["ratio"]]
    np.testing.assert_allclose(r0["t"], full["t"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r0["ratio"], full["ratio"], rtol=1e-4)
    got = torch.cat([r0["sol"], r1["sol"]])
    assert torch.allclose(got, sol[-1], rtol=1e-4, atol=1e-5)


def _train_step(x, traces):
    """one interpolated-adjoint training step (loss = SUM over the batch, so per-shard losses add up)"""
    import torchdyn_b200  # noqa: F401
    from torchdyn_b200.core import NeuralODE
    from cpu_kernel_double import OracleSpline
    om = sys.modules["torchdyn_b200.numerics.odeint"]
    sys.modules["torchdyn_b200.numerics.sensitivity"].NaturalCubicSpline = OracleSpline    # CPU double of the spline kernels
    torch.manual_seed(5)
    # scaled weights + long interval: the controller is ACTIVE (error ratios 0.2-0.9), so the step sequence is decided
    # by the norm and not by fp32 rounding noise of an error estimate far below the tolerance
    net = torch.nn.Sequential(torch.nn.Linear(4, 64), torch.nn.Tanh(), torch.nn.Linear(64, 4))
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(4.0)
    model = NeuralODE(net, solver="tsit5", sensitivity="interpolated_adjoint", atol=1e-4, rtol=1e-4)
    x = x.clone().requires_grad_(True)
    om.TRACE_SINK = traces
    try:
        _, sol = model(x, torch.tensor([0.0, 4.0]))
        sol[-1].pow(2).sum().backward()
    finally:
        om.TRACE_SINK = None
    return net, x.grad


def _worker_interp(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from torchdyn_b200 import _kernels, distributed
    from cpu_kernel_double import CpuKernels
    k = CpuKernels()
    _kernels.set_test_backend(lambda x: k)
    distributed.enable()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(77, 4, generator=g) * 2           # unequal shards (39 + 38)
    shard = distributed.shard(x, rank, world).contiguous()
    traces = []
    net, gx = _train_step(shard, traces)
    distributed.allreduce_gradients(net.parameters())
    torch.save(dict(traces=[dict(t=t["t"], dt=t["dt"], ratio=t["ratio"], dt0=t["dt0"]) for t in traces], gx=gx,
                    gp=[p.grad.clone() for p in net.parameters()]), os.path.join(out_dir, f"interp{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_interpolated_adjoint_matches_the_unsharded_run(tmp_path):
    """mu (the parameter-gradient block) sits inside the error norm of the interpolated adjoint (reference
    sensitivity.py:152): its stage derivatives are all-reduced and it is counted once, so the sharded backward takes the
    unsharded run's steps and -- after allreduce_gradients -- yields its gradients."""
    sys.path.insert(0, HERE)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_interp, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "interp0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "interp1.pt", weights_only=False)
    assert len(r0["traces"]) == len(r1["traces"]) == 2                 # forward solve + one backward interval
    for a, b in zip(r0["traces"], r1["traces"]):
        assert a["t"] == b["t"] and a["dt"] == b["dt"] and a["ratio"] == b["ratio"] and a["dt0"] == b["dt0"]
    for a, b in zip(r0["gp"], r1["gp"]):
        assert torch.equal(a, b)                                       # after the all-reduce every rank holds the sum

    from torchdyn_b200 import _kernels
    from cpu_kernel_double import CpuKernels
    torch.set_num_threads(1)
    k = CpuKernels()
    _kernels.set_test_backend(lambda x: k)
    try:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(77, 4, generator=g) * 2
        traces = []
        net, gx = _train_step(x, traces)
    finally:
        _kernels.set_test_backend(None)
    # forward solve: active controller -> the unsharded run's decisions (CPU sgemm rounds differently for 39- and
    # 77-row batches, f differs in the last bits and dt inherits ~1e-4)
    full, sh = traces[0], r0["traces"][0]
    assert len(full["t"]) == len(sh["t"]) and [r <= 1 for r in sh["ratio"]] == full["accept"]
    np.testing.assert_allclose(sh["dt0"], full["dt0"], rtol=1e-5)
    np.testing.assert_allclose(sh["t"], full["t"], rtol=1e-3, atol=1e-7)
    # backward interval: same initial step; its first steps grow out of a tiny dt0 with error ratios ~1e-4, i.e. at the
    # fp32 noise floor of the embedded estimate, where the order in which d(mu)/dt is summed over the batch (per-rank
    # partials + all-reduce vs one sum) moves dt by percent: the step count may differ by a step or two
    full, sh = traces[1], r0["traces"][1]
    np.testing.assert_allclose(sh["dt0"], full["dt0"], rtol=1e-5)
    assert abs(len(full["t"]) - len(sh["t"])) <= 2
    scale = float(gx.abs().max())
    assert torch.allclose(torch.cat([r0["gx"], r1["gx"]]), gx, rtol=2e-3, atol=2e-3 * scale)
    for a, p in zip(r0["gp"], net.parameters()):
        assert torch.allclose(a, p.grad, rtol=2e-3, atol=2e-3 * float(p.grad.abs().max()))


def test_shard_helper_covers_the_batch():
    from torchdyn_b200 import distributed
    x = torch.arange(10).reshape(10, 1)
    parts = [distributed.shard(x, r, 4) for r in range(4)]
    assert [p.shape[0] for p in parts] == [3, 3, 2, 2] and torch.equal(torch.cat(parts), x)
