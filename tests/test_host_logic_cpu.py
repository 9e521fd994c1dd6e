"""CPU tests of the PRODUCT's host logic (loops, controller, checkpoints, adjoint plumbing, API contract).

The CUDA kernels are replaced by tests/cpu_kernel_double.py (checker arithmetic) via the package's test hook,
so what is exercised here is everything ABOVE the C ABI.  The kernels themselves are checked on the GPU
(tests/test_gpu_parity.py).  Expected values: tests/golden (executed reference)."""
import io
import warnings

import numpy as np
import pytest
import torch
import torch.nn as nn

import torchdyn_b200
from torchdyn_b200 import _kernels
import torchdyn_b200.numerics.sensitivity as sens_mod
from conftest import FUNCTIONAL_CASES, NEURALODE_CASES, VDP, VDP_CASES, load_golden
from cpu_kernel_double import CpuKernels, OracleSpline
from parity_helpers import (check_functional_case, check_neuralode_case, run_cnf_case, run_functional_case,
                            run_neuralode_case)


@pytest.fixture(autouse=True)
def cpu_double(monkeypatch):
    torch.set_num_threads(1)
    k = CpuKernels()
    _kernels.set_test_backend(lambda x: k)
    monkeypatch.setattr(sens_mod, "NaturalCubicSpline", OracleSpline)
    yield k
    _kernels.set_test_backend(None)


@pytest.mark.parametrize("name", FUNCTIONAL_CASES)
def test_functional_odeint_matches_reference(name):
    fx = load_golden(name)
    check_functional_case(fx, *run_functional_case(fx, "cpu"), exact=True)


@pytest.mark.parametrize("name", NEURALODE_CASES)
def test_neuralode_adjoints_match_reference(name):
    fx = load_golden(name)
    check_neuralode_case(fx, run_neuralode_case(fx, "cpu"), exact=True)


@pytest.mark.parametrize("name", VDP_CASES)
def test_vdp_adjoint_matches_reference(name):
    fx = load_golden(name)
    f = VDP(fx["alpha"])
    out = run_neuralode_case(dict(fx, dims=None, act=None, params=None), "cpu", net=f)
    check_neuralode_case(fx, out, exact=True)


def test_cnf_hutchinson_adjoint_matches_reference():
    """cfg 4 (reduced): the CNF vector field (autograd inside f, second-order autograd in the adjoint) through the
    product's loops -- bit-identical to the executed reference with checker arithmetic below the ABI."""
    fx = load_golden("c4r_cnf_hutch")
    out = run_cnf_case(fx, "cpu")
    # the fixture's gradient is w.r.t. the AUGMENTED input
    check_neuralode_case(fx, out, exact=True)


def test_solver_registry_and_errors():
    from torchdyn_b200.numerics import SOLVER_DICT, str_to_solver, odeint
    for name in ("euler", "rk4", "rk-4", "RungeKutta4", "dopri5", "DormandPrince45", "DormandPrince5", "tsit5",
                 "Tsitouras45", "Tsitouras5"):
        s = str_to_solver(name)
        assert s.stepping_class in ("fixed", "adaptive")
    with pytest.raises(KeyError):
        str_to_solver("not-a-solver")
    x = torch.randn(4, 2)
    with pytest.raises(NotImplementedError):
        odeint(lambda t, x: x, x, torch.zeros(2, 3), "dopri5")
    with pytest.warns(UserWarning, match="no effect on fixed-step"):
        odeint(lambda t, x: -x, x, torch.linspace(0, 1, 3), "rk4", atol=1e-5)
    with pytest.warns(UserWarning, match="save_at has no effect"):
        odeint(lambda t, x: -x, x, torch.linspace(0, 1, 3), "dopri5", save_at=torch.tensor([1.0]))
    assert SOLVER_DICT["dopri5"]().order == 5 and SOLVER_DICT["tsit5"]().order == 5 and SOLVER_DICT["rk4"]().order == 4


def test_no_cpu_fallback_without_test_hook():
    _kernels.set_test_backend(None)
    from torchdyn_b200.numerics import odeint
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        odeint(lambda t, x: -x, torch.randn(4, 2), torch.linspace(0, 1, 3), "dopri5")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        odeint(lambda t, x: -x, torch.randn(4, 2), torch.linspace(0, 1, 3), "rk4")


@pytest.mark.parametrize("solver,kw", [("dopri5", dict(atol=1e-4, rtol=1e-4)), ("tsit5", dict(atol=1e-4, rtol=1e-4)),
                                       ("rk4", {}), ("euler", {}), ("dopri5", dict(atol=1e-4, rtol=1e-4, interpolator="4th"))])
def test_sensitivity_autograd_backprop_through_the_solver(solver, kw):
    """sensitivity='autograd' (the NeuralODE default): gradients flow through the differentiable stage combinations.
    Forward values are bit-identical to the oracle; gradients agree to rounding (the reference back-propagates
    through each element-wise op, here through the fused linear map)."""
    from oracle import erk_oracle as eo
    torch.manual_seed(7)
    net = nn.Sequential(nn.Linear(3, 16), nn.Tanh(), nn.Linear(16, 3))
    x0 = torch.randn(12, 3)
    t_span = torch.linspace(0, 1, 4)
    model = torchdyn_b200.NeuralODE(net, solver=solver, sensitivity="autograd", **kw)
    xa = x0.clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, sol = model(xa, t_span)
    sol[-1].pow(2).sum().backward()
    ga, gpa = xa.grad.clone(), [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    ora = eo.OracleNeuralODE(net, solver=solver, sensitivity="autograd", **kw)
    xb = x0.clone().requires_grad_(True)
    _, sol_o = ora(xb, t_span)
    sol_o[-1].pow(2).sum().backward()
    assert torch.equal(sol.detach(), sol_o.detach())
    # dt0 = init_step(x0, f0) carries a (tiny) gradient in the reference that is deliberately dropped here
    assert torch.allclose(ga, xb.grad, rtol=2e-3, atol=1e-5)
    for a, p in zip(gpa, net.parameters()):
        assert torch.allclose(a, p.grad, rtol=2e-3, atol=1e-5)


def test_reversed_span_returns_negated_times():
    from torchdyn_b200.numerics import odeint
    with torch.no_grad():
        t_eval, sol = odeint(lambda t, x: -x, torch.ones(3, 2), torch.tensor([1.0, 0.0]), "dopri5", atol=1e-6, rtol=1e-6)
    assert torch.allclose(t_eval, torch.tensor([-1.0, 0.0]))
    assert torch.allclose(sol[-1], torch.full((3, 2), float(np.e)), rtol=1e-4)


def test_user_solver_plugin_dict_state_and_args():
    """A user subclass overriding `step` runs unchanged (reference test/models/test_ode.py:80-99)."""
    from torchdyn_b200.numerics import Euler

    class DictEuler(Euler):
        def step(self, f, x, t, dt, k1=None, args=None):
            out = {"x": x["x"] + dt * f(t, x["x"]) * args["scale"], "n": x["n"] + 1}
            return None, out, None

    model = torchdyn_b200.NeuralODE(lambda t, x, args={}: -x, solver=DictEuler(), sensitivity="autograd")
    x0 = {"x": torch.ones(2, 3), "n": torch.zeros(1)}
    with torch.no_grad():
        t_eval, sol = model(x0, torch.linspace(0, 1, 11), args={"scale": 1.0})
    assert isinstance(sol, dict) and sol["x"].shape == (11, 2, 3) and sol["n"][-1].item() == 10
    assert torch.allclose(sol["x"][-1], torch.full((2, 3), 0.9 ** 10), rtol=1e-5)


def test_user_adaptive_solver_module_runs_its_own_step():
    from torchdyn_b200.numerics import DormandPrince45, odeint

    class MyDopri(DormandPrince45):
        calls = 0

        def step(self, f, x, t, dt, k1=None, args=None):
            MyDopri.calls += 1
            return super().step(f, x, t, dt, k1=k1, args=args)

    with torch.no_grad():
        t1, s1 = odeint(lambda t, x: -x, torch.ones(4, 2), torch.linspace(0, 1, 3), MyDopri(), atol=1e-5, rtol=1e-5)
        t2, s2 = odeint(lambda t, x: -x, torch.ones(4, 2), torch.linspace(0, 1, 3), "dopri5", atol=1e-5, rtol=1e-5)
    assert MyDopri.calls > 0 and torch.allclose(s1, s2, rtol=1e-6) and torch.equal(t1, t2)


def test_hypersolvers_match_the_oracle_formula():
    """Base step on the kernels + dt**(p+1) * g(t, x) (reference numerics/solvers/hyper.py): against the oracle's base
    step with the same correction added by hand."""
    from oracle import erk_oracle as eo
    from torchdyn_b200.numerics.solvers import HyperEuler, HyperMidpoint, HyperRungeKutta4
    from torchdyn_b200.numerics import odeint
    torch.manual_seed(5)
    f_net = nn.Sequential(nn.Linear(3, 8), nn.Tanh(), nn.Linear(8, 3))
    g_net = nn.Sequential(nn.Linear(3, 8), nn.Tanh(), nn.Linear(8, 3))
    f = lambda t, x: f_net(x)       # noqa: E731
    g = lambda t, x: g_net(x)       # noqa: E731
    x0 = torch.randn(6, 3)
    t_span = torch.linspace(0, 1, 6)
    for cls, name, order in ((HyperEuler, "euler", 1), (HyperMidpoint, "midpoint", 2), (HyperRungeKutta4, "rk4", 4)):
        with torch.no_grad():
            _, sol = odeint(f, x0, t_span, cls(g))
            tab = eo.Tableau(name)
            x, t, dt = x0, t_span[0], t_span[1] - t_span[0]
            for i in range(1, len(t_span)):
                _, xn, _, _ = eo.rk_step(tab, f, x, t, dt)
                x = xn + dt ** (order + 1) * g(t, x)
                t = t + dt
                if i < len(t_span) - 1:
                    dt = t_span[i + 1] - t
        assert torch.allclose(sol[-1], x, rtol=1e-6, atol=1e-7), name


def test_neuralode_api_surface():
    net = nn.Sequential(nn.Linear(2, 8), nn.Tanh(), nn.Linear(8, 2))
    model = torchdyn_b200.NeuralODE(net, solver="dopri5", sensitivity="adjoint", return_t_eval=False)
    assert "NFE" in repr(model) and model.sensalg == "adjoint" and model.vf.nfe == 0
    with torch.no_grad():
        sol = model(torch.randn(5, 2))
        assert sol.shape == (2, 5, 2)                      # default t_span = linspace(0, 1, 2), sol only
        traj = model.trajectory(torch.randn(5, 2), torch.linspace(0, 1, 30))
    assert traj.shape == (30, 5, 2)
    assert model.vf.nfe > 0
    # pickling contract (reference test/test_problem.py:20-32)
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    with torch.no_grad():
        assert again(torch.zeros(3, 2)).shape == (2, 3, 2)
    prob = torchdyn_b200.ODEProblem(net, "tsit5", sensitivity="interpolated_adjoint")
    assert prob.atol == 1e-4 and prob.atol_adjoint == 1e-6
    with pytest.raises(KeyError):
        torchdyn_b200.NeuralODE(net, solver="dopri5", solver_adjoint=nn.Identity())   # must be a NAME (reference quirk)


def test_save_at_subset_fixed_step():
    from torchdyn_b200.numerics import odeint
    t_span = torch.linspace(0, 1, 11)
    with torch.no_grad():
        t_eval, sol = odeint(lambda t, x: -x, torch.ones(2, 2), t_span, "rk4", save_at=t_span[[0, 5, 10]])
    assert sol.shape == (3, 2, 2) and torch.allclose(sol[-1], torch.full((2, 2), float(np.exp(-1.0))), rtol=1e-5)
    with pytest.raises(AssertionError):
        odeint(lambda t, x: -x, torch.ones(2, 2), t_span, "rk4", save_at=torch.tensor([0.123]))


def test_nan_error_ratio_raises_instead_of_hanging():
    from torchdyn_b200.numerics import odeint
    with torch.no_grad(), pytest.raises(RuntimeError, match="did not terminate"):
        odeint(lambda t, x: x * float("nan"), torch.ones(2, 2), torch.linspace(0, 1, 2), "dopri5")


class _PosField(torch.nn.Module):
    order = 1

    def __init__(self, params):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(3, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        with torch.no_grad():
            for p, v in zip(self.net.parameters(), params):
                p.copy_(v)

    def forward(self, t, x):
        return self.net(x)


def test_odeint_symplectic_alf_bit_identical_to_reference():
    """Fixed-step asynchronous leapfrog through odeint_symplectic (reference odeint.py:95-153, ode.py:107-135): the
    executed-reference fixture, forward and reversed span, plus the reference's error behaviour."""
    from torchdyn_b200.numerics import AsynchronousLeapfrog, odeint_symplectic
    fx = load_golden("fn_alf_symplectic")
    f = _PosField(fx["params"])
    with torch.no_grad():
        t1, s1 = odeint_symplectic(f, fx["x"], fx["t_span"], "alf")
        t2, s2 = odeint_symplectic(f, fx["x"], torch.linspace(1, 0, 5), AsynchronousLeapfrog())
    assert torch.equal(s1, fx["sol"]) and torch.equal(t1, fx["t_eval"])
    assert torch.equal(s2, fx["sol_rev"]) and torch.equal(t2, fx["t_rev"])
    with pytest.raises(RuntimeError, match="order"):
        odeint_symplectic(lambda t, x: x, fx["x"], fx["t_span"], "alf")
    f.order = 2
    with pytest.raises(RuntimeError, match="first-order symplectic"):
        odeint_symplectic(f, fx["x"], fx["t_span"], "alf")
    f.order = 1
    with pytest.raises(NotImplementedError):
        odeint_symplectic(f, fx["x"], fx["t_span"], AsynchronousLeapfrog(stepping_class="adaptive"))


class _Ball(torch.nn.Module):
    def forward(self, t, x):
        return torch.cat([x[..., 1:], -9.81 * torch.ones_like(x[..., :1])], -1)


class _Bounce:
    def check_event(self, t, x):
        return bool((x[..., 0] <= 0).any())

    def jump_map(self, t, x):
        return torch.cat([x[..., :1] * 0 + 1e-3, -0.8 * x[..., 1:]], -1)


@pytest.mark.parametrize("solver", ["dopri5", "tsit5"])
def test_odeint_hybrid_bouncing_ball_bit_identical_to_reference(solver):
    """Event detection, bisection of the triggering step and the jump map (reference odeint.py:183-323)."""
    from torchdyn_b200.numerics import odeint_hybrid
    fx = load_golden("fn_hybrid_bouncing_ball")
    with torch.no_grad():
        t_eval, sol = odeint_hybrid(_Ball(), fx["x"], fx["t_span"], fx["j_span"], solver, [_Bounce()], atol=fx["atol"],
                                    rtol=fx["rtol"], event_tol=fx["event_tol"])
    assert torch.equal(t_eval, fx[solver]["t_eval"]) and torch.equal(sol, fx[solver]["sol"])


@pytest.mark.parametrize("sens", ["adjoint", "interpolated_adjoint", "autograd"])
def test_odeproblem_optimizable_params_matches_reference(sens):
    """ODEProblem on a parameter-free callable whose trainable tensor is handed over through `optimizable_params`
    (reference core/problems.py:82-93): time-dependent field, gradients w.r.t. the state and that tensor."""
    from torchdyn_b200.core import ODEProblem
    fx = load_golden("api_optimizable_params")
    A = nn.Parameter(fx["A"].clone())
    prob = ODEProblem(lambda t, x, args={}: torch.tanh(x @ A.T) * (1 + t), solver="dopri5", sensitivity=sens, atol=1e-5,
                      rtol=1e-5, atol_adjoint=1e-5, rtol_adjoint=1e-5, optimizable_params=[A])
    x = fx["x"].clone().requires_grad_(True)
    _, sol = prob.odeint(x, fx["t_span"])
    sol[-1].pow(2).mean().backward()
    ref = fx["ref"][sens]
    assert torch.equal(sol.detach(), ref["sol"])
    if sens == "autograd":      # differently shaped autograd graph: fp32 round-off (see parity_helpers.check_neuralode_case)
        assert torch.allclose(x.grad, ref["gx"], rtol=2e-6, atol=1e-7) and torch.allclose(A.grad, ref["gA"], rtol=2e-6, atol=2e-7)
    else:
        assert torch.equal(x.grad, ref["gx"]) and torch.equal(A.grad, ref["gA"])
