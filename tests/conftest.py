import os
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# ---- strict / loose bookkeeping of the end-to-end fixture comparisons (tests/parity_helpers.py::assert_traces) ----------
# Every fixture comparison records, per odeint call, whether the STRICT step-sequence comparison passed.  The table is
# printed at the end of the pytest run (so GPUTEST logs show it) and every GPU fixture test asserts that a fixture whose
# name is not in tests/golden/LOOSE_ALLOWED.json went strict on every call.
STRICT_REPORT = {}        # fixture name -> [bool per odeint call]


def loose_allowed():
    import json
    with open(os.path.join(GOLDEN, "LOOSE_ALLOWED.json")) as fh:
        return set(json.load(fh)["fixtures"])


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if not STRICT_REPORT:
        return
    tr = terminalreporter
    tr.write_sep("-", "fixture step-sequence comparison: strict (S) / loose fallback (L) per odeint call")
    allowed = loose_allowed()
    for name in sorted(STRICT_REPORT):
        flags = STRICT_REPORT[name]
        mark = "strict" if all(flags) else ("LOOSE (allow-listed)" if name in allowed else "LOOSE (NOT allowed)")
        tr.write_line(f"  {name:40s} {''.join('S' if f else 'L' for f in flags):24s} {mark}")
    n_loose = sum(not all(f) for f in STRICT_REPORT.values())
    tr.write_line(f"  {len(STRICT_REPORT) - n_loose} of {len(STRICT_REPORT)} fixtures strict on every call")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def build_mlp(dims, act, params=None, device="cpu"):
    acts = {"tanh": nn.Tanh, "softplus": nn.Softplus, "relu": nn.ReLU}
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if i < len(dims) - 2:
            layers.append(acts[act]())
    net = nn.Sequential(*layers)
    if params is not None:
        with torch.no_grad():
            for p, v in zip(net.parameters(), params):
                p.copy_(v)
    return net.to(device)


class VDP(nn.Module):
    """Van der Pol oscillator used by the vdp_* fixtures (same as oracle/make_golden.py)."""

    def __init__(self, alpha):
        super().__init__()
        self.alpha = alpha
        self.p = nn.Parameter(torch.tensor([1.0]))

    def forward(self, t, x, args={}):
        x1, x2 = x[..., :1], x[..., 1:2]
        return torch.cat([x2 * self.p, self.alpha * (1 - x1 ** 2) * x2 - x1], -1)


def special_net(kind, params=None, device="cpu"):
    """The non-MLP vector fields of the fixtures (same construction as oracle/make_golden.py, with this package's layers)."""
    from torchdyn_b200.nn import DataControl, DepthCat
    if kind == "order2":
        net = nn.Sequential(nn.Linear(4, 16), nn.Tanh(), nn.Linear(16, 2))
    elif kind == "depthcat_datacontrol":
        net = nn.Sequential(DataControl(), DepthCat(1), nn.Linear(7, 16), nn.Tanh(), nn.Linear(16, 3))
    else:
        raise ValueError(kind)
    if params is not None:
        with torch.no_grad():
            for p, v in zip(net.parameters(), params):
                p.copy_(v)
    return net.to(device)


INTEGRALS = {"sq": lambda t, x: 0.5 * x.pow(2).sum(1)}          # running cost of the integral-loss fixtures

LOSSES = {
    "il_aug": lambda sol, y: sol[-1][:, 1:].pow(2).mean() + sol[-1][:, 0].mean(),
    "sq": lambda sol, y: sol[-1].pow(2).sum(),
    "sqmean": lambda sol, y: sol[-1].pow(2).mean(),
    "ce": lambda sol, y: nn.CrossEntropyLoss()(sol[-1], y),
}

NEURALODE_CASES = ["c1_moons_dopri5_adjoint", "c1_moons_tsit5_interp", "mlp3_tsit5_adjoint_tspan5",
                   "mlp_relu_dopri5_interp_tspan4", "c3r_tsit5_interp", "act_dopri5_adjoint", "act_tsit5_interp",
                   "w256_tsit5_interp", "w256_dopri5_adjoint", "act_w256x2_dopri5_adjoint", "act_w1024_tsit5_interp",
                   "il_dopri5_adjoint", "il_tsit5_interp", "il_rk4_autograd",
                   "ho_order2_dopri5_adjoint", "dc_depthcat_tsit5_autograd", "dc_depthcat_dopri5_interp"]
FUNCTIONAL_CASES = ["fn_euler_tspan10", "fn_midpoint_tspan10", "fn_rk4_tspan10", "fn_dopri5_tspan10", "fn_tsit5_tspan10",
                    "fn_dopri5_4th_tspan10", "fn_dopri5_reversed", "fn_tsit5_all_eval", "fn_rk4_save_at", "c2r_rk4_100",
                    "act_tsit5_fn", "act_dopri5_4th_fn", "act_w256x2_tsit5_fn", "act_w256x2_dopri5_fn"]
# fixtures whose controller is ACTIVE (error ratios ~0.2-0.9, rejections): the GPU must reproduce the step sequence
ACTIVE_CASES = ["act_dopri5_adjoint", "act_tsit5_interp", "act_tsit5_fn", "act_dopri5_4th_fn",
                "act_w256x2_dopri5_adjoint", "act_w1024_tsit5_interp"]
# value tolerance of the active fixtures on the GPU (expanding dynamics amplify the last bits of f; default 2e-4)
ACTIVE_STATE_RTOL = {"act_w256x2_dopri5_adjoint": 1e-4, "act_w1024_tsit5_interp": 1e-3}
VDP_CASES = ["vdp_dopri5_adjoint", "vdp_tsit5_adjoint", "vdp_dopri5_interpolated_adjoint",
             "vdp_tsit5_interpolated_adjoint"]
