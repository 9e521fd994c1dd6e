"""Shared comparison code: run the PRODUCT (torchdyn_b200) on a fixture and compare with what the executed
reference produced (tests/golden) -- used by the CPU host-logic tests (kernel double) and the GPU parity tests
(real kernels, through the C ABI)."""
import warnings

import numpy as np
import torch

import torchdyn_b200
from torchdyn_b200.numerics import odeint as product_odeint
import sys
odeint_mod = sys.modules["torchdyn_b200.numerics.odeint"]   # the package re-exports a function of the same name
from conftest import LOSSES, VDP, build_mlp

# Tolerances (north_star): identical accepted-step count / accept-reject sequence; t, dt sequence equal to fp32
# round-off (CPU vs GPU powf / tanhf / GEMM summation order differ in the last bits); final state and gradients
# within rtol 1e-5 (plus an absolute floor of 1e-6 * scale for entries near zero).
T_RTOL = 2e-6
STATE_RTOL = 1e-5


def assert_close(a, b, what, rtol=STATE_RTOL, atol_scale=2e-6):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    atol = atol_scale * max(float(b.abs().max()), 1e-30)
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), f"{what}: max abs diff {float((a - b).abs().max()):.3e} (atol {atol:.2e}, rtol {rtol})"


def assert_traces(ref_traces, got_traces, ratio_rtol=5e-2, exact=False):
    assert len(ref_traces) == len(got_traces), (len(ref_traces), len(got_traces))
    for r, g in zip(ref_traces, got_traces):
        if exact:
            assert r["dt0"] == g["dt0"] and r["t"] == g["t"] and r["dt"] == g["dt"] and r["ratio"] == g["ratio"]
            continue
        assert len(r["t"]) == len(g["t"]), f"attempted-step count differs: ref {len(r['t'])} vs {len(g['t'])}"
        ref_accept = [x <= 1 for x in r["ratio"]]
        assert ref_accept == g["accept"], "accept/reject sequence differs"
        np.testing.assert_allclose(g["dt0"], r["dt0"], rtol=1e-5)
        np.testing.assert_allclose(g["t"], r["t"], rtol=T_RTOL, atol=1e-7)
        np.testing.assert_allclose(g["dt"], r["dt"], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(g["ratio"], r["ratio"], rtol=ratio_rtol)


def run_neuralode_case(fx, device, net=None):
    net = build_mlp(fx["dims"], fx["act"], fx["params"], device) if net is None else net
    y = fx["y"].to(device) if fx.get("y") is not None else None
    model = torchdyn_b200.NeuralODE(net, **fx["kw"])
    x = fx["x"].clone().to(device).requires_grad_(True)
    odeint_mod.TRACE_SINK = traces = []
    try:
        t_eval, sol = model(x, fx["t_span"])
        nfe_f = model.vf.nfe
        loss = LOSSES[fx["loss"]](sol, y) if "loss" in fx else sol[-1].sum()
        loss.backward()
    finally:
        odeint_mod.TRACE_SINK = None
    return dict(model=model, net=net, x=x, t_eval=t_eval, sol=sol, loss=loss, traces=traces, nfe_fwd=nfe_f)


def check_neuralode_case(fx, out, exact=False):
    """exact=True (CPU host-logic tests, checker arithmetic below the ABI): bit-identical to the executed reference."""
    ref = fx["ref"]
    assert_traces(ref["traces"], out["traces"], exact=exact)
    if exact:
        assert torch.equal(out["t_eval"].detach(), ref["t_eval"]) and torch.equal(out["sol"][-1].detach(), ref["sol_last"])
        assert torch.equal(out["x"].grad, ref["gx"])
        assert all(torch.equal(p.grad, g) for p, g in zip(out["net"].parameters(), ref["gp"]))
    assert out["sol"].shape[0] == ref["sol_len"]
    np.testing.assert_allclose(out["t_eval"].detach().cpu().numpy(), ref["t_eval"].numpy(), rtol=T_RTOL, atol=1e-7)
    assert_close(out["sol"][-1], ref["sol_last"], "final state")
    assert_close(out["loss"], ref["loss"], "loss")
    assert_close(out["x"].grad, ref["gx"], "dL/dx0", rtol=2e-5, atol_scale=5e-6)
    for i, (p, g) in enumerate(zip(out["net"].parameters(), ref["gp"])):
        assert_close(p.grad, g, f"dL/dtheta[{i}]", rtol=2e-5, atol_scale=5e-6)
    assert out["nfe_fwd"] == ref["nfe_fwd"] and out["model"].vf.nfe == ref["nfe_total"], \
        (out["nfe_fwd"], ref["nfe_fwd"], out["model"].vf.nfe, ref["nfe_total"])


def run_functional_case(fx, device, f=None):
    net = build_mlp(fx["dims"], fx["act"], fx["params"], device)
    kw = dict(fx["kw"])
    if "save_at" in kw:
        kw["save_at"] = kw["save_at"].to(device)
    solver = kw.pop("solver")
    odeint_mod.TRACE_SINK = traces = []
    try:
        with torch.no_grad(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t_eval, sol = product_odeint(f if f is not None else (lambda t, x: net(x)), fx["x"].to(device), fx["t_span"],
                                         solver, **kw)
    finally:
        odeint_mod.TRACE_SINK = None
    return t_eval, sol, traces


def check_functional_case(fx, t_eval, sol, traces, exact=False):
    ref = fx["ref"]
    if exact:
        assert torch.equal(t_eval.detach().cpu(), ref["t_eval"]) and torch.equal(sol.cpu(), ref["sol"])
        assert_traces(ref["traces"], traces, exact=True)
    assert sol.shape == ref["sol"].shape, (sol.shape, ref["sol"].shape)
    np.testing.assert_allclose(t_eval.detach().cpu().numpy(), ref["t_eval"].numpy(), rtol=T_RTOL, atol=1e-7)
    assert_close(sol, ref["sol"], "solution")
    if ref["traces"]:
        assert_traces(ref["traces"], traces)
