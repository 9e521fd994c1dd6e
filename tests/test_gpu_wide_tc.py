"""GPU tests of the layer-wise tensor-core kernels for wide MLP vector fields (tdyn_lin_tc_apply, tdyn_wgrad_tc)
and of the adjoint assembled from them (fused.FusedWideAdjoint), against fp64 torch references.

The split-precision scheme (TF32 main term + bf16 correction chain) keeps fp32-accurate products; the tensor core
accumulates with truncation, so the bound is a few 1e-6 of the output scale (same bound as the fused 32-256-256-32
stage kernel, test_gpu_parity.py; measured ~1e-6, benchmarks/wide_tc_accuracy.py)."""
import numpy as np
import pytest
import torch

from torchdyn_b200 import _kernels, fused
from torchdyn_b200.numerics import odeint
from torchdyn_b200._lib import ACT_RELU, ACT_SOFTPLUS, ACT_TANH
from conftest import build_mlp
from parity_helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ACTS = {"tanh": ACT_TANH, "softplus": ACT_SOFTPLUS, "relu": ACT_RELU}


@pytest.fixture(autouse=True)
def real_backend():
    _kernels.set_test_backend(None)
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _act64(name, z):
    return {"tanh": torch.tanh, "softplus": torch.nn.functional.softplus, "relu": torch.relu}[name](z)


def _dact64_from_out(name, a):
    if name == "tanh":
        return 1 - a * a
    if name == "softplus":
        return -torch.expm1(-a)
    return (a > 0).to(a.dtype)


LIN_SHAPES = [(1024, 128, 300), (128, 1024, 257), (32, 256, 129), (256, 32, 128), (256, 256, 500), (64, 64, 200),
              (64, 32, 1), (128, 128, 20000), (2048, 32, 64)]


@pytest.mark.parametrize("n_out,k_in,rows", LIN_SHAPES)
@pytest.mark.parametrize("act", ["tanh", "softplus", "relu"])
def test_lin_tc_forward_modes(n_out, k_in, rows, act):
    """mode 0 (linear, bias, scale -1) and mode 1 (activation) of one layer against fp64."""
    kern = _kernels.get(torch.empty(1, device=DEV))
    assert kern.lin_tc_supported(n_out, k_in)
    torch.manual_seed(n_out + 3 * k_in + rows)
    w = (torch.randn(n_out, k_in, device=DEV) / k_in ** 0.5).contiguous()
    b = torch.randn(n_out, device=DEV)
    x = torch.randn(rows, k_in, device=DEV)
    img = kern.lin_tc_prepare(w, False)
    z64 = x.double() @ w.double().T + b.double()
    out = torch.full((rows, n_out), float("nan"), device=DEV)
    kern.lin_tc_apply(img, n_out, k_in, x, b, None, out, ACTS[act], 0, -1.0, rows)
    assert_close(out, -z64, f"linear {n_out}x{k_in}", rtol=5e-6, atol_scale=5e-6)
    out2 = torch.full((rows, n_out), float("nan"), device=DEV)
    kern.lin_tc_apply(img, n_out, k_in, x, b, None, out2, ACTS[act], 1, 1.0, rows)
    assert_close(out2, _act64(act, z64), f"{act} {n_out}x{k_in}", rtol=5e-6, atol_scale=1e-5)
    again = torch.empty_like(out2)
    kern.lin_tc_apply(img, n_out, k_in, x, b, None, again, ACTS[act], 1, 1.0, rows)
    assert torch.equal(out2, again)


@pytest.mark.parametrize("n_out,k_in,rows", LIN_SHAPES)
@pytest.mark.parametrize("act", ["tanh", "softplus", "relu"])
def test_lin_tc_input_gradient(n_out, k_in, rows, act):
    """mode 2 with the transposed image: gin[r, k] = (sum_n g[r, n] W[n, k]) * act'(a[r, k]) (no bias)."""
    kern = _kernels.get(torch.empty(1, device=DEV))
    torch.manual_seed(7 + n_out + 3 * k_in + rows)
    w = (torch.randn(n_out, k_in, device=DEV) / n_out ** 0.5).contiguous()      # layer: k_in -> n_out
    g = torch.randn(rows, n_out, device=DEV)
    a = _act64(act, torch.randn(rows, k_in, device=DEV).double()).float()      # saved activation outputs of the layer below
    if not kern.lin_tc_supported(k_in, n_out):
        pytest.skip("transposed shape not in the supported class")
    img_t = kern.lin_tc_prepare(w, True)
    want = (g.double() @ w.double()) * _dact64_from_out(act, a.double())
    out = torch.full((rows, k_in), float("nan"), device=DEV)
    kern.lin_tc_apply(img_t, k_in, n_out, g, None, a, out, ACTS[act], 2, 1.0, rows)
    assert_close(out, want, f"dgrad {act} {n_out}x{k_in}", rtol=5e-6, atol_scale=5e-6)


WGRAD_SHAPES = [(1024, 128, 3000), (128, 1024, 1000), (32, 256, 777), (256, 32, 300), (256, 256, 5000), (64, 64, 100),
                (128, 128, 1), (128, 128, 40000), (256, 64, 256), (2048, 32, 513)]


@pytest.mark.parametrize("n_out,k_in,rows", WGRAD_SHAPES)
def test_wgrad_tc_matches_fp64(n_out, k_in, rows):
    kern = _kernels.get(torch.empty(1, device=DEV))
    torch.manual_seed(n_out + k_in + rows)
    g = torch.randn(rows, n_out, device=DEV)
    x = torch.randn(rows, k_in, device=DEV) + 0.25          # non-zero mean: the sums do not cancel
    dw = torch.full((n_out, k_in), float("nan"), device=DEV)
    db = torch.full((n_out,), float("nan"), device=DEV)
    ws = kern.wgrad_tc_workspace(n_out, k_in, rows)
    kern.wgrad_tc(g, n_out, x, k_in, dw, db, -1.0, ws, rows)
    want_w = -(g.double().T @ x.double())
    want_b = -g.double().sum(0)
    assert_close(dw, want_w, f"dW {n_out}x{k_in} rows {rows}", rtol=1e-5, atol_scale=5e-6)
    assert_close(db, want_b, f"db {n_out} rows {rows}", rtol=1e-5, atol_scale=2e-6)
    dw2, db2 = torch.empty_like(dw), torch.empty_like(db)
    kern.wgrad_tc(g, n_out, x, k_in, dw2, db2, -1.0, ws, rows)
    assert torch.equal(dw, dw2) and torch.equal(db, db2), "parameter-gradient reduction must be deterministic"


WIDE_NETS = [([128, 1024, 128], "tanh", 300), ([32, 256, 256, 32], "tanh", 200), ([32, 256, 256, 32], "softplus", 1000),
             ([64, 512, 256, 64], "relu", 129), ([32, 256, 32], "tanh", 4099), ([128, 1024, 1024, 128], "softplus", 64),
             ([32, 128, 256, 128, 32], "tanh", 515), ([64, 2048, 64], "relu", 77)]


@pytest.mark.parametrize("dims,act,rows", WIDE_NETS)
def test_wide_adjoint_matches_autograd(dims, act, rows):
    """dx = f(x), dlam = -lam^T df/dx, dmu = -sum lam^T df/dtheta (flat, parameters() order) against fp64 autograd."""
    torch.manual_seed(rows + sum(dims))
    net = build_mlp(dims, act).to(DEV)
    fm = fused.FusedMLP(net)
    kern = _kernels.get(torch.empty(1, device=DEV))
    assert fm.ready_wide(kern)
    D = dims[0]
    x = torch.randn(rows, D, device=DEV)
    lam = torch.randn(rows, D, device=DEV)
    net64 = build_mlp(dims, act, [p.detach() for p in net.parameters()], DEV).double()
    x64 = x.double().requires_grad_(True)
    f64 = net64(x64)
    grads = torch.autograd.grad(f64, (x64,) + tuple(net64.parameters()), -lam.double())
    want_dlam, want_dmu = grads[0], torch.cat([g.flatten() for g in grads[1:]])
    adj = fused.FusedWideAdjoint(fm, [], kern)
    for sign in (1.0, -1.0):
        dx, dlam = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
        dmu = torch.full((fm.n_params,), float("nan"), device=DEV)
        adj(x.flatten(), lam.flatten(), dx.view(-1), dlam.view(-1), dmu, rows, sign)
        assert_close(dx, sign * f64.detach(), "dx", rtol=5e-6, atol_scale=5e-6)
        assert_close(dlam, sign * want_dlam, "dlam", rtol=1e-5, atol_scale=5e-6)
        assert_close(dmu, sign * want_dmu, "dmu", rtol=2e-5, atol_scale=5e-6)
    # f is optional (interpolated adjoint)
    dlam2, dmu2 = torch.empty_like(x), torch.empty_like(dmu)
    adj(x.flatten(), lam.flatten(), None, dlam2.view(-1), dmu2, rows, -1.0)
    assert torch.equal(dlam2, dlam) and torch.equal(dmu2, dmu)


def test_wide_forward_field_matches_torch():
    """FusedField of a 128-1024-128 net evaluates through the layer-wise kernels (both signs)."""
    torch.manual_seed(0)
    net = build_mlp([128, 1024, 128], "tanh").to(DEV)
    x = torch.randn(1000, 128, device=DEV)
    field = fused.wrap(fused.MLPField(net), x)
    assert isinstance(field, fused.FusedField) and "tdyn_lin_tc_apply" in fused.last_used()
    net64 = build_mlp([128, 1024, 128], "tanh", [p.detach() for p in net.parameters()], DEV).double()   # separate copy
    with torch.no_grad():
        want = net64(x.double())
    field.tdyn_eval(np.float32(0.0), x)                     # builds the weight images (4 launches, once per update)
    n0 = _kernels.launches()
    out = field.tdyn_eval(np.float32(0.0), x)
    assert _kernels.launches() - n0 == 2
    assert_close(out, want, "wide field", rtol=5e-6, atol_scale=5e-6)
    assert_close(field.tdyn_eval(np.float32(0.0), x, -1.0), -want, "wide field, reversed", rtol=5e-6, atol_scale=5e-6)


@pytest.mark.parametrize("sens,solver", [("interpolated_adjoint", "tsit5"), ("adjoint", "dopri5")])
@pytest.mark.parametrize("dims", [[32, 256, 32], [32, 256, 256, 32]])
def test_wide_neuralode_training_step_matches_generic_path(sens, solver, dims):
    """A training step (forward + adjoint) on the tensor-core kernels against the same step with f / VJP as torch
    calls (fused.DISABLED): same accepted-step count, final state and gradients to the split-precision bound."""
    from torchdyn_b200.core import NeuralODE
    torch.manual_seed(3)
    net = build_mlp(dims, "tanh").to(DEV)
    x = torch.randn(512, dims[0], device=DEV)
    t_span = torch.linspace(0, 1, 2)
    res = {}
    for disabled in (True, False):
        fused.DISABLED = disabled
        try:
            model = NeuralODE(net, solver=solver, sensitivity=sens, atol=1e-4, rtol=1e-4, atol_adjoint=1e-4, rtol_adjoint=1e-4)
            net.zero_grad()
            n0 = _kernels.launches()
            _, sol = model(x, t_span)
            sol[-1].pow(2).mean().backward()
            res[disabled] = (sol[-1].detach().clone(), torch.cat([p.grad.flatten() for p in net.parameters()]).clone(),
                             float(model.vf.nfe), _kernels.launches() - n0)
        finally:
            fused.DISABLED = False
    assert res[False][2] == res[True][2], "NFE (step sequence) differs between the fused and the generic path"
    assert_close(res[False][0], res[True][0], "final state", rtol=2e-5, atol_scale=1e-5)
    # (two fp32-accurate evaluations of f and its VJP -- kernel errors ~1e-6 of scale, benchmarks/wide_tc_accuracy.py --
    # pushed through an adaptive solve and its adjoint)
    assert_close(res[False][1], res[True][1], "gradients", rtol=2e-4, atol_scale=1e-4)


@pytest.mark.parametrize("name", ["w256_tsit5_interp", "w256_dopri5_adjoint", "act_w256x2_dopri5_adjoint", "act_w1024_tsit5_interp"])
def test_wide_fixture_runs_on_the_tensor_core_kernels(name):
    """The executed-reference fixtures with 256-wide nets (test_gpu_parity checks their values) really take the
    layer-wise tensor-core path: forward and both halves of the VJP are launches of this library."""
    from conftest import load_golden
    from parity_helpers import check_neuralode_case, run_neuralode_case
    fx = load_golden(name)
    timer = _kernels.KernelTimer()
    _kernels.set_timer(timer)
    try:
        out = run_neuralode_case(fx, DEV)
    finally:
        _kernels.set_timer(None)
    from conftest import ACTIVE_STATE_RTOL
    check_neuralode_case(fx, out, state_rtol=ACTIVE_STATE_RTOL.get(name, 1e-5), name=name)
    names = {r[0] for r in timer.records}
    if name == "act_w256x2_dopri5_adjoint":
        # the 32-256-256-32 class: forward stages, the adjoint's forward half (activations saved) and its input-gradient
        # chain all run on the fused stage kernel; the parameter gradients on tdyn_wgrad_tc
        assert "mlp_tc_stage" in names and "wgrad_tc" in names and "lin_tc" not in names, names
    else:
        assert "lin_tc" in names and "wgrad_tc" in names, names
    nfe_adj = out["model"].vf.nfe - out["nfe_fwd"]
    n_layers = len(fx["dims"]) - 1
    assert sum(r[0] == "wgrad_tc" for r in timer.records) == n_layers * nfe_adj or fx["kw"]["sensitivity"] == "adjoint"


def test_cfg3_full_size_duplicated_batch_property():
    """BASELINE config 3 at full size on the tensor-core kernels (128-1024-128, 262 144 rows, tsit5 + interpolated
    adjoint).  Size-independent property: for the batch [x0; x0] every global RMS norm equals that of x0 alone (fp64 sums
    double exactly), so the forward controller takes the same steps and both halves of the solution equal the half-batch
    solution BIT FOR BIT (rows are computed independently of their position).  In the backward the parameter-gradient
    block mu sits inside the interpolated adjoint's error norm and doubles with the batch, so the backward steps may
    differ: the two halves of dL/dx0 are bit-identical to each other and agree with the half-batch run to the solver
    tolerance, the parameter gradient of the SUM loss doubles to the same tolerance."""
    from torchdyn_b200.core import NeuralODE
    torch.manual_seed(11)
    net = build_mlp([128, 1024, 128], "tanh").to(DEV)
    x0 = torch.randn(131072, 128, device=DEV)
    t_span = torch.linspace(0, 1, 2)

    def step(x):
        net.zero_grad()
        model = NeuralODE(net, solver="tsit5", sensitivity="interpolated_adjoint")
        xx = x.clone().requires_grad_(True)
        _, sol = model(xx, t_span)
        sol[-1].pow(2).sum().backward()
        assert "tdyn_lin_tc_apply" in fused.last_used()
        return sol[-1].detach(), xx.grad.clone(), torch.cat([p.grad.flatten() for p in net.parameters()]).clone(), float(model.vf.nfe)

    s1, gx1, gp1, nfe1 = step(x0)
    s2, gx2, gp2, nfe2 = step(torch.cat([x0, x0]))
    assert torch.equal(s2[:131072], s1) and torch.equal(s2[131072:], s1)
    assert torch.equal(gx2[:131072], gx2[131072:])
    assert abs(nfe1 - nfe2) <= 12
    assert_close(gx2[:131072], gx1, "dL/dx0 of the duplicated batch", rtol=2e-2, atol_scale=2e-3)
    assert_close(gp2, 2 * gp1, "parameter gradient of the duplicated batch", rtol=2e-2, atol_scale=2e-3)


@pytest.mark.parametrize("case", ["reversed", "dense4th", "all_eval", "rk4_save_at"])
def test_wide_field_through_odeint_variants(case):
    """The functional API on a 64-256-64 field (layer-wise tensor-core forward) against the same call with f as a
    torch call: reversed span, dopri5 with the 4th-order dense output, return_all_eval, fixed-step rk4 with save_at."""
    torch.manual_seed(21)
    net = build_mlp([64, 256, 64], "tanh").to(DEV)
    x = torch.randn(700, 64, device=DEV)
    field = fused.MLPField(net)
    kw = {"reversed": dict(t_span=torch.linspace(1, 0, 4), solver="dopri5", atol=1e-5, rtol=1e-5),
          "dense4th": dict(t_span=torch.linspace(0, 1, 7), solver="dopri5", atol=1e-5, rtol=1e-5, interpolator="4th"),
          "all_eval": dict(t_span=torch.linspace(0, 1, 3), solver="tsit5", atol=1e-5, rtol=1e-5, return_all_eval=True),
          "rk4_save_at": dict(t_span=torch.linspace(0, 1, 21), solver="rk4", save_at=torch.linspace(0, 1, 21)[::5])}[case]
    t_span = kw.pop("t_span")
    solver = kw.pop("solver")
    res = {}
    for disabled in (True, False):
        fused.DISABLED = disabled
        try:
            with torch.no_grad():
                res[disabled] = odeint(field, x, t_span, solver, **kw)
            if not disabled:
                assert "tdyn_lin_tc_apply" in (fused.last_used() or ""), fused.last_used()
        finally:
            fused.DISABLED = False
    (t0, s0), (t1, s1) = res[True], res[False]
    if case == "all_eval":
        # every accepted step is returned: the step SIZES of a default-initialised net at tol 1e-5 are decided at the
        # fp32 noise floor of the error estimate, so only the end points are comparable between two f implementations
        assert abs(t0.shape[0] - t1.shape[0]) <= 2 and torch.allclose(t0[[0, -1]], t1[[0, -1]])
        t0, s0, t1, s1 = t0[-1:], s0[-1:], t1[-1:], s1[-1:]
    assert t0.shape == t1.shape and torch.allclose(t0, t1, rtol=2e-4, atol=1e-6)
    assert_close(s1, s0, f"solution ({case})", rtol=1e-4, atol_scale=2e-5)
