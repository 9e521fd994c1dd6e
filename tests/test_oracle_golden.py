"""Pin the CPU oracle (oracle/erk_oracle.py) against the golden fixtures produced by EXECUTING the reference
(oracle/make_golden.py).  Bit-exact on CPU: t_eval, states, gradients, NFE and the per-attempt (t, dt, ratio)
trace.  When the reference tree is present (build container) also re-run it live."""
import pytest
import torch

from conftest import (FUNCTIONAL_CASES, INTEGRALS, LOSSES, NEURALODE_CASES, VDP, VDP_CASES, build_mlp, load_golden)
from oracle import erk_oracle as eo
from oracle import ref_shim


def _run_oracle(net, fx, loss_fn):
    kw = dict(fx["kw"])
    if fx.get("integral"):
        kw["integral_loss"] = INTEGRALS[fx["integral"]]
    model = eo.OracleNeuralODE(net, **kw)
    x = fx["x"].clone().requires_grad_(True)
    t_eval, sol = model(x, fx["t_span"])
    loss = loss_fn(sol)
    loss.backward()
    return model, x, t_eval, sol, loss


def _check_traces(ref_traces, ora_traces):
    ora_traces = [o for o in ora_traces if o.t]          # (a fixed-step solve records no attempts)
    assert len(ref_traces) == len(ora_traces)
    for r, o in zip(ref_traces, ora_traces):
        assert r["t"] == o.t and r["dt"] == o.dt and r["ratio"] == o.ratio
        assert r["dt0"] == o.dt0


@pytest.mark.parametrize("name", NEURALODE_CASES)
def test_oracle_neuralode_bit_exact(name):
    torch.set_num_threads(1)
    fx = load_golden(name)
    if fx.get("special"):
        pytest.skip("DEFunc features (order, DepthCat, DataControl) belong to the module API, which the oracle does not restate")
    net = build_mlp(fx["dims"], fx["act"], fx["params"])
    model, x, t_eval, sol, loss = _run_oracle(net, fx, lambda s: LOSSES[fx["loss"]](s, fx["y"]))
    ref = fx["ref"]
    assert torch.equal(t_eval, ref["t_eval"]) and torch.equal(sol[-1].detach(), ref["sol_last"])
    assert torch.equal(loss.detach(), ref["loss"]) and torch.equal(x.grad, ref["gx"])
    for p, g in zip(net.parameters(), ref["gp"]):
        assert torch.equal(p.grad, g)
    assert model.nfe == ref["nfe_total"]
    _check_traces(ref["traces"], model.traces)


@pytest.mark.parametrize("name", FUNCTIONAL_CASES)
def test_oracle_functional_bit_exact(name):
    torch.set_num_threads(1)
    fx = load_golden(name)
    net = build_mlp(fx["dims"], fx["act"], fx["params"])
    kw = dict(fx["kw"]); solver = kw.pop("solver")
    tr = eo.Trace()
    with torch.no_grad():
        t_eval, sol = eo.odeint(lambda t, x: net(x), fx["x"], fx["t_span"], solver, trace=tr, **kw)
    assert torch.equal(t_eval, fx["ref"]["t_eval"]) and torch.equal(sol, fx["ref"]["sol"])
    if fx["ref"]["traces"]:
        r = fx["ref"]["traces"][0]
        assert r["t"] == tr.t and r["ratio"] == tr.ratio and r["dt0"] == tr.dt0


@pytest.mark.parametrize("name", VDP_CASES)
def test_oracle_vdp_adjoint_bit_exact(name):
    torch.set_num_threads(1)
    fx = load_golden(name)
    f = VDP(fx["alpha"])
    model, x, t_eval, sol, loss = _run_oracle(f, fx, lambda s: s[-1].sum())
    ref = fx["ref"]
    assert torch.equal(t_eval, ref["t_eval"]) and torch.equal(x.grad, ref["gx"])
    assert torch.equal(f.p.grad, ref["gp"][0]) and model.nfe == ref["nfe_total"]
    _check_traces(ref["traces"], model.traces)


def test_oracle_spline_fixture():
    fx = load_golden("spline_natural")
    co = eo.natural_cubic_coeffs(fx["x"], fx["t"])
    assert torch.equal(co, fx["coeffs"])
    sp = eo.NaturalCubicSpline.from_coeffs(co, fx["t"])
    assert torch.equal(torch.stack([sp.evaluate(v) for v in fx["q"]]), fx["vals"])
    # interpolation property at the knots and C2 continuity of the restated spline
    for i, tk in enumerate(fx["t"]):
        assert torch.allclose(sp.evaluate(tk), fx["x"][..., i, :], atol=1e-5)


def test_oracle_linear_ode_known_answer():
    """dx/dt = A x has the closed form expm(A t) x0: dopri5 at tight tolerance must reproduce it (fp64)."""
    torch.manual_seed(0)
    A = torch.tensor([[0.0, 1.0], [-1.0, -0.1]], dtype=torch.float64)
    x0 = torch.randn(16, 2, dtype=torch.float64)
    t_span = torch.linspace(0, 2, 5, dtype=torch.float64)
    _, sol = eo.odeint(lambda t, x: x @ A.T, x0, t_span, "dopri5", atol=1e-10, rtol=1e-10)
    for i, t in enumerate(t_span):
        assert torch.allclose(sol[i], x0 @ torch.linalg.matrix_exp(A * t).T, atol=1e-8, rtol=1e-8)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    """Execute the real reference here and compare with the oracle on a fresh (non-fixture) input."""
    ref_shim.load()
    from torchdyn.numerics import odeint as ref_odeint
    torch.manual_seed(99)
    net = build_mlp([5, 24, 5], "softplus")
    x = torch.randn(17, 5)
    t_span = torch.linspace(0, 1.5, 4)
    f = lambda t, x: net(x)  # noqa: E731
    with torch.no_grad():
        for solver, kw in (("dopri5", dict(atol=1e-5, rtol=1e-5)), ("tsit5", dict(atol=1e-4, rtol=1e-6)),
                           ("dopri5", dict(atol=1e-4, rtol=1e-4, interpolator="4th")), ("rk4", {}), ("euler", {})):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                tr, sr = ref_odeint(f, x, t_span, solver, **kw)
            to, so = eo.odeint(f, x, t_span, solver, **kw)
            assert torch.equal(tr, to) and torch.equal(sr, so), solver
