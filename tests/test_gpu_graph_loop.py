"""GPU tests of the device-side adaptive loop (CUDA-graph conditional WHILE, torchdyn_b200/numerics/graph_loop.py):
an opaque vector field stays a torch call, the whole accept / reject loop is one graph launch, and the step sequence is
the host-driven loop's (which the fixture tests pin to the executed reference)."""
import sys

import numpy as np
import pytest
import torch

import torchdyn_b200
from torchdyn_b200 import _kernels
from torchdyn_b200.numerics import graph_loop, odeint
from conftest import build_mlp, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
om = sys.modules["torchdyn_b200.numerics.odeint"]


@pytest.fixture(autouse=True)
def real_backend():
    _kernels.set_test_backend(None)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    graph_loop.ENABLED = False
    graph_loop.CACHE = False


def _solve(f, x, ts, solver, device_loop, **kw):
    graph_loop.ENABLED = device_loop
    om.TRACE_SINK = tr = []
    try:
        with torch.no_grad():
            t_eval, sol = odeint(f, x, ts, solver, **kw)
    finally:
        om.TRACE_SINK = None
        graph_loop.ENABLED = False
    return t_eval, sol, tr


@pytest.mark.parametrize("solver", ["dopri5", "tsit5"])
@pytest.mark.parametrize("case", ["two_points", "tspan7", "all_eval", "reversed"])
def test_device_loop_reproduces_the_host_loop(solver, case):
    torch.manual_seed(0)
    net = build_mlp([4, 64, 4], "tanh").to(DEV)
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(3.0)                      # active controller: rejections
    f = lambda t, x: net(x) * (1 + 2 * torch.sin(8 * t))  # noqa: E731   (opaque to the library, fast time dependence)
    x = torch.randn(300, 4, device=DEV) * 2
    ts = {"two_points": torch.tensor([0.0, 3.0]), "tspan7": torch.linspace(0, 3, 7), "all_eval": torch.tensor([0.0, 1.5, 3.0]),
          "reversed": torch.linspace(2, 0, 4)}[case]
    kw = dict(atol=1e-4, rtol=1e-4, return_all_eval=case == "all_eval")
    launches0 = _kernels.launches()
    t_h, s_h, tr_h = _solve(f, x, ts, solver, False, **kw)
    t_d, s_d, tr_d = _solve(f, x, ts, solver, True, **kw)
    assert graph_loop.last_status == "ok", graph_loop.last_status
    assert len(tr_h) == len(tr_d) == 1 and len(tr_h[0]["t"]) >= 5
    assert tr_h[0]["accept"] == tr_d[0]["accept"]
    np.testing.assert_array_equal(np.float32(tr_h[0]["t"]), np.float32(tr_d[0]["t"]))
    np.testing.assert_array_equal(np.float32(tr_h[0]["dt"]), np.float32(tr_d[0]["dt"]))
    np.testing.assert_allclose(tr_h[0]["ratio"], tr_d[0]["ratio"], rtol=1e-6)
    assert torch.equal(t_h, t_d) and s_h.shape == s_d.shape
    assert torch.equal(s_h, s_d), float((s_h - s_d).abs().max())
    assert _kernels.launches() > launches0


def test_device_loop_conv_state_and_cache():
    """4-D state, cuDNN vector field (config-5 shape, reduced); the captured loop is re-used by the next call."""
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 3, 3, padding=1), torch.nn.Tanh()).to(DEV)
    f = lambda t, x: net(x)                  # noqa: E731
    x = torch.randn(8, 3, 12, 12, device=DEV)
    ts = torch.linspace(0, 1, 3)
    t_h, s_h, tr_h = _solve(f, x, ts, "dopri5", False, atol=1e-4, rtol=1e-4)
    graph_loop.CACHE = True
    for _ in range(2):
        t_d, s_d, tr_d = _solve(f, x, ts, "dopri5", True, atol=1e-4, rtol=1e-4)
        assert graph_loop.last_status == "ok", graph_loop.last_status
        assert tr_h[0]["accept"] == tr_d[0]["accept"]
        np.testing.assert_allclose(tr_h[0]["t"], tr_d[0]["t"], rtol=1e-6)
        torch.testing.assert_close(s_d, s_h, rtol=1e-5, atol=1e-6)        # (cuDNN may pick another algorithm under capture)
    x2 = x * 0.5
    _, s_h2, _ = _solve(f, x2, ts, "dopri5", False, atol=1e-4, rtol=1e-4)
    _, s_d2, _ = _solve(f, x2, ts, "dopri5", True, atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(s_d2, s_h2, rtol=1e-5, atol=1e-6)


def test_device_loop_neuralode_forward_counts_nfe_and_adjoint_keeps_the_host_loop():
    """NeuralODE with an opaque conv field: the forward solve runs on the device loop (NFE counted as the reference counts
    it), the backsolve adjoint -- whose VJP is a torch-autograd call inside a running backward pass -- keeps the host-driven
    loop; states, gradients and NFE equal the all-host run."""
    res = {}
    for flag in (False, True):
        torch.manual_seed(6)
        net = torch.nn.Sequential(torch.nn.Conv2d(2, 2, 3, padding=1), torch.nn.Tanh()).to(DEV)
        model = torchdyn_b200.NeuralODE(net, solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint")
        x = torch.randn(4, 2, 8, 8, generator=torch.Generator().manual_seed(1)).to(DEV).requires_grad_(True)
        graph_loop.ENABLED = flag
        try:
            _, sol = model(x, torch.linspace(0, 1, 2))
            st_f, nfe_f = graph_loop.last_status, model.vf.nfe
            sol[-1].pow(2).mean().backward()
            st_b = graph_loop.last_status
        finally:
            graph_loop.ENABLED = False
        res[flag] = (sol.detach(), x.grad.clone(), [p.grad.clone() for p in net.parameters()], model.vf.nfe, nfe_f, st_f, st_b)
    assert res[True][5] == "ok" and "host-side times" in res[True][6], res[True][5:]
    assert res[True][3] == res[False][3] and res[True][4] == res[False][4], (res[True][3:5], res[False][3:5])
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(res[True][1], res[False][1], rtol=1e-4, atol=1e-7)
    for a_, b_ in zip(res[True][2], res[False][2]):
        torch.testing.assert_close(a_, b_, rtol=1e-4, atol=1e-6)


def test_device_loop_falls_back_when_f_cannot_be_captured():
    torch.manual_seed(7)
    net = build_mlp([3, 16, 3], "tanh").to(DEV)

    def f(t, x):
        float(x.sum())          # a host synchronisation: illegal under stream capture
        return net(x)
    x = torch.randn(16, 3, device=DEV)
    t_h, s_h, _ = _solve(f, x, torch.linspace(0, 1, 3), "dopri5", False)
    t_d, s_d, _ = _solve(f, x, torch.linspace(0, 1, 3), "dopri5", True)
    assert graph_loop.last_status.startswith("capture failed"), graph_loop.last_status
    assert torch.equal(s_h, s_d)
