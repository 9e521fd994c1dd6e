"""Regenerate tests/golden/*.pt by EXECUTING the reference (build container only).

    python -m oracle.make_golden            # writes tests/golden/
    python -m oracle.make_golden --check    # also asserts the oracle restatement is bit-identical

Every fixture stores the seeded inputs (state, t_span, parameters) together with what the
unmodified reference (/root/reference, imported through oracle/ref_shim.py) produced on CPU in
fp32: t_eval, sol (or the rows the tests need), gradients, NFE and the per-attempt step trace
(t, dt, error_ratio) captured from outside by wrapping the reference's own hooks.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import erk_oracle as eo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def mlp(dims, act="tanh", seed=0, gain=1.0):
    torch.manual_seed(seed)
    acts = {"tanh": nn.Tanh, "softplus": nn.Softplus, "relu": nn.ReLU}
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(acts[act]())
    net = nn.Sequential(*layers)
    if gain != 1.0:   # "active controller" fixtures: larger weights -> error estimates far above the fp32 noise floor
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(gain)
    return net


def moons(n, seed=0):
    """cfg-1 input: the reference's own ToyDataset moons generator (datasets/static_datasets.py:71-95)."""
    ref_shim.load()
    from torchdyn.datasets import ToyDataset
    np.random.seed(seed)
    torch.manual_seed(seed)
    X, y = ToyDataset().generate(n, "moons", noise=0.4)
    return X.float(), y.long()


class VDP(nn.Module):
    """Van der Pol oscillator, same dynamics as the reference test system (numerics/systems.py:32-41)."""
    def __init__(self, alpha):
        super().__init__()
        self.alpha = alpha
        self.p = nn.Parameter(torch.tensor([1.0]))  # so the adjoint has a parameter to differentiate

    def forward(self, t, x, args={}):
        x1, x2 = x[..., :1], x[..., 1:2]
        return torch.cat([x2 * self.p, self.alpha * (1 - x1 ** 2) * x2 - x1], -1)


def _trace_lists(tr):
    return [dict(t0=r["t0"], T=r["T"], dt0=r["dt0"], t=list(r["t"]), dt=list(r["dt"]), ratio=list(r["ratio"]))
            for r in tr.records]


def run_ref_neuralode(net, x, t_span, loss_fn, **kw):
    ref = ref_shim.load()
    from torchdyn.core import NeuralODE
    model = NeuralODE(net, **kw)
    x = x.clone().requires_grad_(True)
    with ref_shim.StepTrace() as tr:
        t_eval, sol = model(x, t_span)
        nfe_f = float(model.vf.nfe)
        loss = loss_fn(sol)
        loss.backward()
    return dict(t_eval=t_eval.detach(), sol_last=sol[-1].detach(), sol_len=int(sol.shape[0]),
                loss=loss.detach(), gx=x.grad.detach().clone(),
                gp=[p.grad.detach().clone() for p in net.parameters()],
                nfe_fwd=nfe_f, nfe_total=float(model.vf.nfe), traces=_trace_lists(tr))


def run_oracle_neuralode(net, x, t_span, loss_fn, **kw):
    for p in net.parameters():
        p.grad = None
    model = eo.OracleNeuralODE(net, **kw)
    x = x.clone().requires_grad_(True)
    t_eval, sol = model(x, t_span)
    nfe_f = model.nfe
    loss = loss_fn(sol)
    loss.backward()
    return dict(t_eval=t_eval.detach(), sol_last=sol[-1].detach(), sol_len=int(sol.shape[0]), loss=loss.detach(),
                gx=x.grad.detach().clone(), gp=[p.grad.detach().clone() for p in net.parameters()],
                nfe_fwd=nfe_f, nfe_total=model.nfe, traces=model.traces)


def check_same(ref_out, ora_out, name):
    for k in ("t_eval", "sol_last", "loss", "gx"):
        assert torch.equal(ref_out[k], ora_out[k]), f"{name}: oracle != reference on {k}: " \
            f"{(ref_out[k] - ora_out[k]).abs().max()}"
    for a, b in zip(ref_out["gp"], ora_out["gp"]):
        assert torch.equal(a, b), f"{name}: param grad mismatch {(a - b).abs().max()}"
    assert ref_out["nfe_total"] == ora_out["nfe_total"], (name, ref_out["nfe_total"], ora_out["nfe_total"])
    # the reference-side recorder only sees ADAPTIVE odeint calls; the oracle also opens a (then empty) trace for a
    # fixed-step solve under sensitivity='autograd'
    ora_traces = [o for o in ora_out["traces"] if len(o.t) > 0]
    assert len(ref_out["traces"]) == len(ora_traces), name
    for r, o in zip(ref_out["traces"], ora_traces):
        assert r["t"] == o.t and r["dt"] == o.dt and r["ratio"] == o.ratio, f"{name}: trace mismatch"
        assert r["dt0"] == o.dt0


def neuralode_cases():
    cases = {}
    # cfg 1 (BASELINE.json configs[0]): moons 1024x2, MLP 2-64-2 tanh, dopri5 1e-4, backsolve adjoint, CE loss
    X, y = moons(1024, 0)
    ce = nn.CrossEntropyLoss()
    cases["c1_moons_dopri5_adjoint"] = dict(
        dims=[2, 64, 2], act="tanh", seed=0, x=X, y=y, t_span=torch.linspace(0, 1, 2), loss="ce",
        kw=dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint"))
    cases["c1_moons_tsit5_interp"] = dict(
        dims=[2, 64, 2], act="tanh", seed=0, x=X[:256], y=y[:256], t_span=torch.linspace(0, 1, 2), loss="ce",
        kw=dict(solver="tsit5", atol=1e-4, rtol=1e-4, sensitivity="interpolated_adjoint"))
    torch.manual_seed(3)
    cases["mlp3_tsit5_adjoint_tspan5"] = dict(
        dims=[4, 32, 32, 4], act="softplus", seed=1, x=torch.randn(64, 4), y=None, t_span=torch.linspace(0, 1, 5),
        loss="sq", kw=dict(solver="tsit5", atol=1e-5, rtol=1e-5, sensitivity="adjoint",
                           atol_adjoint=1e-5, rtol_adjoint=1e-5))
    torch.manual_seed(4)
    cases["mlp_relu_dopri5_interp_tspan4"] = dict(
        dims=[8, 64, 8], act="relu", seed=2, x=torch.randn(128, 8), y=None, t_span=torch.linspace(0, 2, 4),
        loss="sq", kw=dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="interpolated_adjoint"))
    torch.manual_seed(5)
    # cfg 3 reduced (tsit5 + interpolated adjoint, defaults 1e-3), width 128 instead of 1024, B=128
    cases["c3r_tsit5_interp"] = dict(
        dims=[16, 128, 16], act="tanh", seed=3, x=torch.randn(128, 16), y=None, t_span=torch.linspace(0, 1, 2),
        loss="sqmean", kw=dict(solver="tsit5", sensitivity="interpolated_adjoint"))
    # active-controller cases (single interval, scaled weights): error ratios ~0.2-0.9, the step sequence is decided
    # by the controller rather than by checkpoint clamps or rounding noise
    torch.manual_seed(6)
    xa = torch.randn(256, 4) * 2
    cases["act_dopri5_adjoint"] = dict(
        dims=[4, 64, 4], act="tanh", seed=5, gain=3.0, x=xa, y=None, t_span=torch.tensor([0.0, 4.0]), loss="sqmean",
        kw=dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint", atol_adjoint=1e-4, rtol_adjoint=1e-4))
    cases["act_tsit5_interp"] = dict(
        dims=[4, 64, 4], act="tanh", seed=5, gain=4.0, x=xa, y=None, t_span=torch.tensor([0.0, 4.0]), loss="sqmean",
        kw=dict(solver="tsit5", atol=1e-4, rtol=1e-4, sensitivity="interpolated_adjoint"))
    # integral losses (reference sensitivity.py:67-69, 141-143; core/defunc.py:68-77): adjoint / interpolated adjoint add
    # -dg/dx to the co-state dynamics; sensitivity='autograd' propagates the running cost in column 0 of the state
    torch.manual_seed(13)
    xi = torch.randn(96, 4)
    cases["il_dopri5_adjoint"] = dict(
        dims=[4, 32, 4], act="tanh", seed=13, x=xi, y=None, t_span=torch.linspace(0, 1, 3), loss="sqmean", integral="sq",
        kw=dict(solver="dopri5", atol=1e-5, rtol=1e-5, sensitivity="adjoint", atol_adjoint=1e-5, rtol_adjoint=1e-5))
    cases["il_tsit5_interp"] = dict(
        dims=[4, 32, 4], act="tanh", seed=13, x=xi, y=None, t_span=torch.linspace(0, 1, 3), loss="sqmean", integral="sq",
        kw=dict(solver="tsit5", atol=1e-5, rtol=1e-5, sensitivity="interpolated_adjoint"))
    cases["il_rk4_autograd"] = dict(
        dims=[4, 32, 4], act="tanh", seed=13, x=torch.cat([torch.zeros(96, 1), xi], 1), y=None,
        t_span=torch.linspace(0, 1, 11), loss="il_aug", integral="sq", kw=dict(solver="rk4", sensitivity="autograd"))
    # DEFunc features (reference core/defunc.py:42-94, core/neuralde.py:96-121): higher-order dynamics, depth
    # concatenation, data control
    torch.manual_seed(14)
    cases["ho_order2_dopri5_adjoint"] = dict(
        special="order2", dims=None, act=None, seed=14, x=torch.randn(48, 4), y=None, t_span=torch.linspace(0, 1, 3),
        loss="sqmean", kw=dict(order=2, solver="dopri5", atol=1e-5, rtol=1e-5, sensitivity="adjoint", atol_adjoint=1e-5,
                               rtol_adjoint=1e-5))
    cases["dc_depthcat_tsit5_autograd"] = dict(
        special="depthcat_datacontrol", dims=None, act=None, seed=15, x=torch.randn(48, 3), y=None,
        t_span=torch.linspace(0, 1, 3), loss="sqmean", kw=dict(solver="tsit5", atol=1e-5, rtol=1e-5, sensitivity="autograd"))
    cases["dc_depthcat_dopri5_interp"] = dict(
        special="depthcat_datacontrol", dims=None, act=None, seed=15, x=torch.randn(48, 3), y=None,
        t_span=torch.linspace(0, 1, 3), loss="sqmean", kw=dict(solver="dopri5", atol=1e-5, rtol=1e-5,
                                                                sensitivity="interpolated_adjoint"))
    # wide nets (widths multiples of 32, hidden 256): the layer-wise tensor-core kernels (tdyn_lin_tc_apply / tdyn_wgrad_tc)
    torch.manual_seed(9)
    cases["w256_tsit5_interp"] = dict(
        dims=[32, 256, 32], act="tanh", seed=9, x=torch.randn(256, 32), y=None, t_span=torch.linspace(0, 1, 2),
        loss="sqmean", kw=dict(solver="tsit5", atol=1e-4, rtol=1e-4, sensitivity="interpolated_adjoint"))
    torch.manual_seed(10)
    cases["w256_dopri5_adjoint"] = dict(
        dims=[64, 256, 64], act="softplus", seed=10, x=torch.randn(128, 64), y=None, t_span=torch.linspace(0, 1, 2),
        loss="sqmean", kw=dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint", atol_adjoint=1e-4,
                               rtol_adjoint=1e-4))
    # ACTIVE controller on the tensor-core paths: gain-scaled wide nets whose step sequence is decided by the
    # controller (ratios 0.15-0.95, rejections at 1.2-3.8) and NOT by rounding noise: tolerances 1e-2 / 1e-3 keep the
    # embedded error estimates 100-1000x above the fp32 noise of f, the first backward step is clipped at max_factor.
    # (i) the config-2 network class 32-256-256-32 (fused tcgen05 stage kernel forward, layer-wise kernels in the adjoint),
    # (ii) config 3's own shape 128-1024-128 (K = 1024: the accumulator-flush path of tdyn_lin_tc_apply / tdyn_wgrad_tc)
    torch.manual_seed(16)
    cases["act_w256x2_dopri5_adjoint"] = dict(
        dims=[32, 256, 256, 32], act="tanh", seed=16, gain=4.0, x=torch.randn(256, 32), y=None, t_span=torch.tensor([0.0, 1.0]),
        loss="sq", kw=dict(solver="dopri5", atol=1e-2, rtol=1e-2, sensitivity="adjoint", atol_adjoint=1e-2,
                           rtol_adjoint=1e-2))
    torch.manual_seed(17)
    cases["act_w1024_tsit5_interp"] = dict(
        dims=[128, 1024, 128], act="tanh", seed=17, gain=6.0, x=torch.randn(256, 128), y=None, t_span=torch.tensor([0.0, 2.0]),
        loss="sqmean", kw=dict(solver="tsit5", atol=1e-3, rtol=1e-3, sensitivity="interpolated_adjoint"))
    return cases


def special_net(kind, layers, seed):
    """Vector fields that are not plain MLPs; `layers` provides DepthCat / DataControl (reference torchdyn.nn here)."""
    torch.manual_seed(seed)
    if kind == "order2":            # second-order system on [q, p]: the net maps the full state to the acceleration
        return nn.Sequential(nn.Linear(4, 16), nn.Tanh(), nn.Linear(16, 2))
    if kind == "depthcat_datacontrol":   # state 3 + control 3 + depth 1 -> 3
        return nn.Sequential(layers.DataControl(), layers.DepthCat(1), nn.Linear(7, 16), nn.Tanh(), nn.Linear(16, 3))
    raise ValueError(kind)


INTEGRALS = {"sq": lambda t, x: 0.5 * x.pow(2).sum(1)}          # running cost of the integral-loss cases

LOSSES = {
    # terminal cost + the integral propagated in column 0 of the augmented state (sensitivity='autograd' + integral_loss)
    "il_aug": lambda sol, y: sol[-1][:, 1:].pow(2).mean() + sol[-1][:, 0].mean(),
    "sq": lambda sol, y: sol[-1].pow(2).sum(),
    "sqmean": lambda sol, y: sol[-1].pow(2).mean(),
    "ce": lambda sol, y: nn.CrossEntropyLoss()(sol[-1], y),
}


def functional_cases():
    """odeint(f, x, t_span, solver, ...) on an MLP vector field; full solutions stored."""
    cases = {}
    torch.manual_seed(11)
    x = torch.randn(32, 3)
    for solver in ("euler", "midpoint", "rk4", "dopri5", "tsit5"):
        cases[f"fn_{solver}_tspan10"] = dict(dims=[3, 16, 3], act="tanh", seed=7, x=x, t_span=torch.linspace(0, 2, 10),
                                             kw=dict(solver=solver, atol=1e-5, rtol=1e-5))
    cases["fn_dopri5_4th_tspan10"] = dict(dims=[3, 16, 3], act="tanh", seed=7, x=x, t_span=torch.linspace(0, 2, 10),
                                          kw=dict(solver="dopri5", atol=1e-5, rtol=1e-5, interpolator="4th"))
    cases["fn_dopri5_reversed"] = dict(dims=[3, 16, 3], act="tanh", seed=7, x=x, t_span=torch.linspace(1, 0, 4),
                                       kw=dict(solver="dopri5", atol=1e-6, rtol=1e-6))
    cases["fn_tsit5_all_eval"] = dict(dims=[3, 16, 3], act="tanh", seed=7, x=x, t_span=torch.linspace(0, 3, 3),
                                      kw=dict(solver="tsit5", atol=1e-6, rtol=1e-6, return_all_eval=True))
    torch.manual_seed(6)
    xa = torch.randn(256, 4) * 2
    cases["act_tsit5_fn"] = dict(dims=[4, 64, 4], act="tanh", seed=5, gain=4.0, x=xa, t_span=torch.tensor([0.0, 4.0]),
                                 kw=dict(solver="tsit5", atol=1e-4, rtol=1e-4, return_all_eval=True))
    cases["act_dopri5_4th_fn"] = dict(dims=[4, 64, 4], act="tanh", seed=5, gain=4.0, x=xa, t_span=torch.linspace(0, 4, 9),
                                      kw=dict(solver="dopri5", atol=1e-4, rtol=1e-4, interpolator="4th"))
    # ACTIVE controller on the persistent tensor-core integrator (tdyn_mlp_tc_solve_adaptive): the config-2 network class,
    # gain 6 (ratios 0.1-0.7, rejections at 1.03-15), more than two 128-row tiles (ragged last tile), interior output times, all accepted steps returned
    torch.manual_seed(18)
    xw = torch.randn(300, 32)
    cases["act_w256x2_tsit5_fn"] = dict(dims=[32, 256, 256, 32], act="tanh", seed=16, gain=6.0, x=xw,
                                        t_span=torch.tensor([0.0, 1.0, 2.0]),
                                        kw=dict(solver="tsit5", atol=1e-2, rtol=1e-2, return_all_eval=True))
    cases["act_w256x2_dopri5_fn"] = dict(dims=[32, 256, 256, 32], act="tanh", seed=16, gain=6.0, x=torch.cat([xw, xw[:33] * 0.5]),
                                         t_span=torch.linspace(0, 2, 4), kw=dict(solver="dopri5", atol=1e-3, rtol=1e-3))
    cases["fn_rk4_save_at"] = dict(dims=[3, 16, 3], act="tanh", seed=7, x=x, t_span=torch.linspace(0, 1, 21),
                                   kw=dict(solver="rk4"), save_last=True)
    # cfg 2 reduced: rk4 100 steps, MLP 32-256-256-32, B=64, forward only, save last
    torch.manual_seed(12)
    cases["c2r_rk4_100"] = dict(dims=[32, 256, 256, 32], act="tanh", seed=8, x=torch.randn(64, 32),
                                t_span=torch.linspace(0, 1, 101), kw=dict(solver="rk4"), save_last=True)
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated fixture names (default: all)")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    from torchdyn.numerics import odeint as ref_odeint
    torch.set_num_threads(1)  # reduction order of torch CPU kernels is thread-count independent for these sizes, but pin it

    for name, c in neuralode_cases().items():
        if only and name not in only:
            continue
        if c.get("special"):
            import torchdyn.nn as ref_layers
            net = special_net(c["special"], ref_layers, c["seed"])
        else:
            net = mlp(c["dims"], c["act"], c["seed"], c.get("gain", 1.0))
        loss_fn = (lambda sol, c=c: LOSSES[c["loss"]](sol, c["y"]))
        kw_ref = dict(c["kw"])
        if c.get("integral"):
            kw_ref["integral_loss"] = INTEGRALS[c["integral"]]
        out = run_ref_neuralode(net, c["x"], c["t_span"], loss_fn, **kw_ref)
        if args.check:
            check_same(out, run_oracle_neuralode(net, c["x"], c["t_span"], loss_fn, **kw_ref), name)
        fix = dict(kind="neuralode", dims=c["dims"], act=c["act"], x=c["x"], y=c["y"], t_span=c["t_span"], loss=c["loss"],
                   kw=c["kw"], integral=c.get("integral"), special=c.get("special"), params=[p.detach().clone() for p in net.parameters()], ref=out)
        torch.save(fix, os.path.join(OUT, name + ".pt"))
        ntr = [len(r["t"]) for r in out["traces"]]
        print(f"{name}: nfe={out['nfe_total']:.0f} attempts per odeint call={ntr} len(t_eval)={len(out['t_eval'])}")

    for name, c in functional_cases().items():
        if only and name not in only:
            continue
        net = mlp(c["dims"], c["act"], c["seed"], c.get("gain", 1.0))
        kw = dict(c["kw"])
        if c.get("save_last"):
            kw["save_at"] = c["t_span"][-1:]
        f = lambda t, x: net(x)
        with torch.no_grad(), ref_shim.StepTrace() as tr:
            t_eval, sol = ref_odeint(f, c["x"], c["t_span"], **kw)
        if args.check:
            okw = {k: v for k, v in kw.items()}
            solver = okw.pop("solver")
            otr = eo.Trace()
            with torch.no_grad():
                t2, s2 = eo.odeint(f, c["x"], c["t_span"], solver, trace=otr, **okw)
            assert torch.equal(t_eval, t2) and torch.equal(sol, s2), f"{name}: oracle != reference"
            if tr.records:
                assert tr.records[0]["t"] == otr.t and tr.records[0]["ratio"] == otr.ratio, name
        fix = dict(kind="functional", dims=c["dims"], act=c["act"], x=c["x"], t_span=c["t_span"], kw=kw,
                   params=[p.detach().clone() for p in net.parameters()],
                   ref=dict(t_eval=t_eval, sol=sol, traces=_trace_lists(tr)))
        torch.save(fix, os.path.join(OUT, name + ".pt"))
        print(f"{name}: sol {tuple(sol.shape)} attempts={[len(r['t']) for r in tr.records]}")

    if only:
        return
    # Van der Pol adjoint (reduced test/test_sensitivity.py:40-76): gradient wrt x0, t_span of 20 points
    for sens in ("adjoint", "interpolated_adjoint"):
        for solver in ("dopri5", "tsit5"):
            name = f"vdp_{solver}_{sens}"
            torch.manual_seed(21)
            x = torch.randn(64, 2)
            f = VDP(0.5)
            t_span = torch.linspace(0, 1, 20)
            kw = dict(solver=solver, atol=1e-4, rtol=1e-4, atol_adjoint=1e-4, rtol_adjoint=1e-4, sensitivity=sens)
            out = run_ref_neuralode(f, x, t_span, lambda sol: sol[-1].sum(), **kw)
            if args.check:
                check_same(out, run_oracle_neuralode(f, x, t_span, lambda sol: sol[-1].sum(), **kw), name)
            torch.save(dict(kind="vdp", alpha=0.5, x=x, t_span=t_span, kw=kw, ref=out), os.path.join(OUT, name + ".pt"))
            print(f"{name}: nfe={out['nfe_total']:.0f} calls={len(out['traces'])}")

    # cfg 4 reduced: CNF with Hutchinson trace (torchdyn/models/cnf.py:26-65), MLP 2-64-64-64-2 softplus, dopri5 1e-4,
    # backsolve adjoint, B = 256, fixed noise (the reference samples it once per forward, core/neuralde.py:112-116)
    from torchdyn.models import CNF, hutch_trace
    from torchdyn.nn import Augmenter
    torch.manual_seed(41)
    xc = torch.randn(256, 2)
    noise = torch.randn(256, 2)

    class FixedNoise:
        def __init__(self, v):
            self.v = v

        def sample(self, shape):
            return self.v

    netc = mlp([2, 64, 64, 64, 2], "softplus", 9)
    cnf = CNF(netc, trace_estimator=hutch_trace, noise_dist=FixedNoise(noise))
    kwc = dict(solver="dopri5", atol=1e-4, rtol=1e-4, sensitivity="adjoint")

    def cnf_loss(sol):
        prior = torch.distributions.MultivariateNormal(torch.zeros(2), torch.eye(2))
        return -(prior.log_prob(sol[-1][:, 1:]) - sol[-1][:, 0]).mean()

    xa0 = Augmenter(1, 1)(xc)
    outc = run_ref_neuralode(cnf, xa0, torch.linspace(0, 1, 2), cnf_loss, **kwc)
    if args.check:
        cnf.noise = noise
        check_same(outc, run_oracle_neuralode(cnf, xa0, torch.linspace(0, 1, 2), cnf_loss, **kwc), "c4r_cnf_hutch")
    torch.save(dict(kind="cnf", dims=[2, 64, 64, 64, 2], act="softplus", x=xc, noise=noise, t_span=torch.linspace(0, 1, 2),
                    kw=kwc, params=[p.detach().clone() for p in netc.parameters()], ref=outc),
               os.path.join(OUT, "c4r_cnf_hutch.pt"))
    print(f"c4r_cnf_hutch: nfe={outc['nfe_total']:.0f} attempts={[len(r['t']) for r in outc['traces']]}")

    # fixed-step asynchronous leapfrog through odeint_symplectic (reference odeint.py:95-153, ode.py:107-135), forward and
    # reversed span; the vector field carries `order = 1` (v' = f(t, x) on the position half of the state)
    from torchdyn.numerics import odeint_symplectic as ref_symplectic

    class _PosField(nn.Module):
        order = 1

        def __init__(self):
            super().__init__()
            torch.manual_seed(0)
            self.net = nn.Sequential(nn.Linear(3, 16), nn.Tanh(), nn.Linear(16, 3))

        def forward(self, t, x):
            return self.net(x)

    fa = _PosField()
    torch.manual_seed(1)
    xs = torch.randn(7, 6)
    with torch.no_grad():
        t_a, s_a = ref_symplectic(fa, xs, torch.linspace(0, 1, 11), 'alf')
        t_b, s_b = ref_symplectic(fa, xs, torch.linspace(1, 0, 5), 'alf')
    torch.save(dict(x=xs, t_span=torch.linspace(0, 1, 11), params=[p.detach().clone() for p in fa.parameters()], sol=s_a,
                    t_eval=t_a, sol_rev=s_b, t_rev=t_b), os.path.join(OUT, "fn_alf_symplectic.pt"))
    print(f"fn_alf_symplectic: sol {tuple(s_a.shape)}")

    # hybrid system (reference odeint.py:183-323): bouncing ball [height, velocity], jump = restitution 0.8 at the floor
    from torchdyn.numerics import odeint_hybrid as ref_hybrid

    class _Ball(nn.Module):
        def forward(self, t, x):
            return torch.cat([x[..., 1:], -9.81 * torch.ones_like(x[..., :1])], -1)

    class _Bounce:
        def check_event(self, t, x):
            return bool((x[..., 0] <= 0).any())

        def jump_map(self, t, x):
            return torch.cat([x[..., :1] * 0 + 1e-3, -0.8 * x[..., 1:]], -1)

    hyb = dict(x=torch.tensor([[1.0, 0.0]]), t_span=torch.linspace(0, 2, 5), j_span=3, atol=1e-5, rtol=1e-5, event_tol=1e-4)
    with torch.no_grad():
        for sv in ("dopri5", "tsit5"):
            t_h, s_h = ref_hybrid(_Ball(), hyb["x"], hyb["t_span"], hyb["j_span"], sv, [_Bounce()], atol=hyb["atol"],
                                  rtol=hyb["rtol"], event_tol=hyb["event_tol"])
            hyb[sv] = dict(t_eval=t_h, sol=s_h)
    torch.save(hyb, os.path.join(OUT, "fn_hybrid_bouncing_ball.pt"))
    print(f"fn_hybrid_bouncing_ball: {[tuple(hyb[sv]['sol'].shape) for sv in ('dopri5', 'tsit5')]}")

    # ODEProblem on a parameter-free callable with `optimizable_params` (reference core/problems.py:82-93)
    from torchdyn.core import ODEProblem as RefProblem
    opt = {}
    for sens in ("adjoint", "interpolated_adjoint", "autograd"):
        torch.manual_seed(0)
        A = nn.Parameter(torch.randn(3, 3) * 0.5)
        prob = RefProblem(lambda t, x, args={}: torch.tanh(x @ A.T) * (1 + t), solver="dopri5", sensitivity=sens, atol=1e-5,
                          rtol=1e-5, atol_adjoint=1e-5, rtol_adjoint=1e-5, optimizable_params=[A])
        xo = torch.randn(20, 3, generator=torch.Generator().manual_seed(1)).requires_grad_(True)
        _, so = prob.odeint(xo, torch.linspace(0, 1, 3))
        so[-1].pow(2).mean().backward()
        opt[sens] = dict(sol=so.detach(), gx=xo.grad.clone(), gA=A.grad.clone())
    torch.manual_seed(0)
    torch.save(dict(A=torch.randn(3, 3) * 0.5, x=torch.randn(20, 3, generator=torch.Generator().manual_seed(1)),
                    t_span=torch.linspace(0, 1, 3), ref=opt), os.path.join(OUT, "api_optimizable_params.pt"))
    print("api_optimizable_params: ok")

    # spline: self-consistency vector (UNPINNED against torchcde, see oracle/__init__.py)
    torch.manual_seed(31)
    xs = torch.randn(5, 9, 3)
    ts = torch.sort(torch.rand(9))[0]
    co = eo.natural_cubic_coeffs(xs, ts)
    sp = eo.NaturalCubicSpline.from_coeffs(co, ts)
    q = torch.tensor([-0.1, float(ts[0]), float(ts[3]), 0.5 * float(ts[3] + ts[4]), float(ts[-1]), 1.5])
    torch.save(dict(kind="spline", x=xs, t=ts, coeffs=co, q=q, vals=torch.stack([sp.evaluate(v) for v in q])),
               os.path.join(OUT, "spline_natural.pt"))
    print("done ->", OUT)


if __name__ == "__main__":
    main()
