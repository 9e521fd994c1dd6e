"""torch-CPU restatement of torchdyn's batched explicit-RK hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this file (see oracle/__init__.py).  It is a *restatement*, not a copy: the
algorithm is table-driven here, but every arithmetic expression keeps the
reference's association order because PyTorch rounds each elementwise op
separately and the accepted-step sequence depends on those bits
(SURVEY.md Appendix A).  Citations are relative to /root/reference.

Parity: PINNED against the executed reference through tests/golden/*.pt
(bit-exact on CPU), except the natural cubic spline (third-party torchcde ^0.2.3,
absent from the reference tree; restated from its published algorithm: UNPINNED).
"""
from __future__ import annotations

from fractions import Fraction as Fr
from typing import Callable, List, Optional

import torch
from torch import Tensor

# --------------------------------------------------------------------------------------
# Butcher tableaus.  Reference: torchdyn/numerics/solvers/_constants.py:20-46 (rk4, dopri5),
# :49-112 (tsit5), :121-127 (4th-order dense-output b_mid).  Values are python doubles
# rounded to the tensor dtype by torch.tensor; b_err is stored as (b_sol - b_err) with the
# subtraction done IN THE TENSOR DTYPE (:46, :112).
# --------------------------------------------------------------------------------------

def _f(x) -> float:
    return float(x.numerator) / float(x.denominator) if isinstance(x, Fr) else float(x)


_DOPRI5 = dict(
    c=[Fr(1, 5), Fr(3, 10), Fr(4, 5), Fr(8, 9), 1.0, 1.0],
    a=[[Fr(1, 5)],
       [Fr(3, 40), Fr(9, 40)],
       [Fr(44, 45), Fr(-56, 15), Fr(32, 9)],
       [Fr(19372, 6561), Fr(-25360, 2187), Fr(64448, 6561), Fr(-212, 729)],
       [Fr(9017, 3168), Fr(-355, 33), Fr(46732, 5247), Fr(49, 176), Fr(-5103, 18656)],
       [Fr(35, 384), 0, Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84)]],
    bsol=[Fr(35, 384), 0, Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84), 0],
    bhat=[Fr(1951, 21600), 0, Fr(22642, 50085), Fr(451, 720), Fr(-12231, 42400), Fr(649, 6300), Fr(1, 60)],
)

# Tsitouras 5(4) coefficients (Tsitouras 2011), as decimal literals.
_TSIT5 = dict(
    c=[0.161, 0.327, 0.9, 0.9800255409045096857298102862870245954942137979563024768854764293221195950761080302604, 1.0, 1.0],
    a=[[0.161],
       [-0.8480655492356988544426874250230774675121177393430391537369234245294192976164141156943e-2,
        0.3354806554923569885444268742502307746751211773934303915373692342452941929761641411569],
       [2.897153057105493432130432594192938764924887287701866490314866693455023795137503079289,
        -6.359448489975074843148159912383825625952700647415626703305928850207288721235210244366,
        4.362295432869581411017727318190886861027813359713760212991062156752264926097707165077],
       [5.325864828439256604428877920840511317836476253097040101202360397727981648835607691791,
        -11.74888356406282787774717033978577296188744178259862899288666928009020615663593781589,
        7.495539342889836208304604784564358155658679161518186721010132816213648793440552049753,
        -0.9249506636175524925650207933207191611349983406029535244034750452930469056411389539635e-1],
       [5.861455442946420028659251486982647890394337666164814434818157239052507339770711679748,
        -12.92096931784710929170611868178335939541780751955743459166312250439928519268343184452,
        8.159367898576158643180400794539253485181918321135053305748355423955009222648673734986,
        -0.7158497328140099722453054252582973869127213147363544882721139659546372402303777878835e-1,
        -0.2826905039406838290900305721271224146717633626879770007617876201276764571291579142206e-1],
       [0.9646076681806522951816731316512876333711995238157997181903319145764851595234062815396e-1,
        0.01,
        0.4798896504144995747752495322905965199130404621990332488332634944254542060153074523509,
        1.379008574103741893192274821856872770756462643091360525934940067397245698027561293331,
        -3.290069515436080679901047585711363850115683290894936158531296799594813811049925401677,
        2.324710524099773982415355918398765796109060233222962411944060046314465391054716027841]],
    bhat=[0.9468075576583945807478876255758922856117527357724631226139574065785592789071067303271e-1,
          0.9183565540343253096776363936645313759813746240984095238905939532922955247253608687270e-2,
          0.4877705284247615707855642599631228241516691959761363774365216240304071651579571959813,
          1.234297566930478985655109673884237654035539930748192848315425833500484878378061439761,
          -2.707712349983525454881109975059321670689605166938197378763992255714444407154902012702,
          1.866628418170587035753719399566211498666255505244122593996591602841258328965767580089,
          1.0 / 66.0],
)
_TSIT5["bsol"] = list(_TSIT5["a"][5]) + [0.0]

_RK4 = dict(c=[0.0, 0.5, 0.5, 1.0], a=[[0.5], [0.0, 0.5], [0.0, 0.0, 1.0]],
            bsol=[Fr(1, 6), Fr(1, 3), Fr(1, 3), Fr(1, 6)])

_BMID_4TH = [0.10013431883002395, 0, 0.3918321794184259, -0.02982460176594817,
             0.05893268337240795, -0.04497888809104361, 0.023904308236133973]

# name -> (canonical, order, stepping class).  Reference registry: solvers/ode.py:320-325.
_ALIASES = {
    "euler": "euler", "midpoint": "midpoint",
    "rk4": "rk4", "rk-4": "rk4", "RungeKutta4": "rk4",
    "dopri5": "dopri5", "DormandPrince45": "dopri5", "DormandPrince5": "dopri5",
    "tsit5": "tsit5", "Tsitouras45": "tsit5", "Tsitouras5": "tsit5",
}
_ORDER = {"euler": 1, "midpoint": 2, "rk4": 4, "dopri5": 5, "tsit5": 5}
_ADAPTIVE = {"euler": False, "midpoint": False, "rk4": False, "dopri5": True, "tsit5": True}


class Tableau:
    def __init__(self, name: str, dtype=torch.float32):
        self.name = _ALIASES[name]  # KeyError on unknown names, like SOLVER_DICT[...] (ode.py:335)
        self.order = _ORDER[self.name]
        self.adaptive = _ADAPTIVE[self.name]
        spec = {"dopri5": _DOPRI5, "tsit5": _TSIT5, "rk4": _RK4, "euler": None, "midpoint": None}[self.name]
        self.dtype = dtype
        if spec is None:
            self.c = self.a = self.bsol = self.berr = None
            return
        tt = lambda row: torch.tensor([_f(v) for v in row], dtype=dtype)
        self.c = tt(spec["c"])
        self.a = [tt(r) for r in spec["a"]]
        self.bsol = tt(spec["bsol"])
        self.berr = (self.bsol - tt(spec["bhat"])) if "bhat" in spec else None

    def cast(self, like: Tensor):
        """templates.py:30-40 sync_device_dtype: tableau is CAST (not rebuilt) to x's dtype."""
        if self.c is not None:
            self.c = self.c.to(like)
            self.a = [r.to(like) for r in self.a]
            self.bsol = self.bsol.to(like)
            if self.berr is not None:
                self.berr = self.berr.to(like)
        self.dtype = like.dtype
        return self


def _chain(x: Tensor, dt, coefs, ks) -> Tensor:
    """``x + dt*a0*k0 + dt*a1*k1 + ...`` -- each term is ((dt*a)*k), summed left to right.
    Shape of stages k2, k4..k7 of dopri5/tsit5 (solvers/ode.py:148,150-153 and :170,172-175)."""
    y = x
    for a, k in zip(coefs, ks):
        y = y + (dt * a) * k
    return y


def _inner(x: Optional[Tensor], dt, coefs, ks) -> Tensor:
    """``x + dt*(a0*k0 + a1*k1 + ...)`` (or ``dt*(...)`` when x is None).
    Shape of k3, x_sol and err of dopri5/tsit5 (ode.py:149,154,155) and of every rk4 stage (:100-103)."""
    s = coefs[0] * ks[0]
    for a, k in zip(coefs[1:], ks[1:]):
        s = s + a * k
    return dt * s if x is None else x + dt * s


def rk_step(tab: Tableau, f: Callable, x: Tensor, t, dt, k1: Optional[Tensor] = None):
    """One explicit RK step.  Returns (f_new, x_new, err, stages) for adaptive tableaus
    and (None, x_new, None, None) for fixed ones, like the reference steppers
    (ode.py:68-71 euler, :97-104 rk4, :145-156 dopri5, :167-178 tsit5)."""
    if k1 is None:
        k1 = f(t, x)
    if tab.name == "euler":
        return None, x + dt * k1, None, None
    if tab.name == "midpoint":                                   # ode.py:75-86
        x_mid = x + 0.5 * dt * k1
        return None, x + dt * f(t + 0.5 * dt, x_mid), None, None
    c, a, bs = tab.c, tab.a, tab.bsol
    if tab.name == "rk4":
        k2 = f(t + c[0] * dt, _inner(x, dt, [a[0]], [k1]))
        k3 = f(t + c[1] * dt, _inner(x, dt, [a[1][0], a[1][1]], [k1, k2]))
        k4 = f(t + c[2] * dt, _inner(x, dt, [a[2][0], a[2][1], a[2][2]], [k1, k2, k3]))
        return None, _inner(x, dt, [bs[0], bs[1], bs[2], bs[3]], [k1, k2, k3, k4]), None, None
    # 7-stage FSAL pairs.  dopri5 multiplies by the 1-element tensor a[0], tsit5 by the scalar a[0][0];
    # both are one fp32 product (ode.py:148 vs :170).
    ks: List[Tensor] = [k1]
    for i in range(6):
        row = [a[i][j] for j in range(i + 1)]
        if i == 0:
            y = _chain(x, dt, [a[0] if tab.name == "dopri5" else a[0][0]], ks)
        elif i == 1:
            y = _inner(x, dt, row, ks)
        else:
            y = _chain(x, dt, row, ks)
        ks.append(f(t + c[i] * dt, y))
    x_new = _inner(x, dt, [bs[j] for j in range(6)], ks[:6])
    err = _inner(None, dt, [tab.berr[j] for j in range(7)], ks)
    return ks[6], x_new, err, tuple(ks)


# --------------------------------------------------------------------------------------
# Norm, initial step, controller.  Reference: torchdyn/numerics/utils.py:33-63.
# --------------------------------------------------------------------------------------

def hairer_norm(v: Tensor) -> Tensor:
    """sqrt(mean(|v|^2)) over ALL elements (utils.py:33-34)."""
    return v.abs().pow(2).mean().sqrt()


def init_step(f, f0, x0, t0, order, atol, rtol):
    """Hairer's starting step (utils.py:37-54).  NB: odeint passes the UN-negated f together
    with the negated t0 on reversed spans (odeint.py:87-89) -- callers here do the same."""
    scale = atol + torch.abs(x0) * rtol
    d0 = hairer_norm(x0 / scale)
    d1 = hairer_norm(f0 / scale)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = torch.tensor(1e-6, dtype=t0.dtype, device=t0.device)
    else:
        h0 = 0.01 * d0 / d1
    f1 = f(t0 + h0, x0 + h0 * f0)
    d2 = hairer_norm((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6, dtype=t0.dtype, device=t0.device), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1).to(t0)


_SAFETY = torch.tensor([0.9])
_MIN_FACTOR = torch.tensor([0.2])
_MAX_FACTOR = torch.tensor([10])  # int64 on purpose: templates.py:25 builds it from the python int 10


@torch.no_grad()
def adapt_step(dt, ratio, order):
    """Step-size controller (utils.py:57-63).  ratio==0 -> x10; ratio<1 lifts min_factor to 1."""
    # (the reference moves these constants to the state's device in sync_device_dtype, templates.py:32-42)
    dev = dt.device
    if ratio == 0:
        return dt * _MAX_FACTOR.to(dev)
    lo = torch.ones_like(dt) if ratio < 1 else _MIN_FACTOR.to(dev)
    expo = torch.tensor(order, dtype=dt.dtype, device=dev).reciprocal()
    return dt * torch.min(_MAX_FACTOR.to(dev), torch.max(_SAFETY.to(dev) / ratio ** expo, lo))


# --------------------------------------------------------------------------------------
# Dense output, reference: torchdyn/numerics/interpolators.py:29-37 (evaluate), :57-63 (fit),
# odeint.py:381 (x_mid).
# --------------------------------------------------------------------------------------

def bmid_4th(dtype=torch.float32) -> Tensor:
    return torch.tensor(_BMID_4TH, dtype=dtype)


def dense_fit_4th(dt, f0, f1, x0, x1, x_mid):
    c1 = 2 * dt * (f1 - f0) - 8 * (x1 + x0) + 16 * x_mid
    c2 = dt * (5 * f0 - 3 * f1) + 18 * x0 + 14 * x1 - 32 * x_mid
    c3 = dt * (f1 - 4 * f0) - 11 * x0 - 5 * x1 + 16 * x_mid
    return [x0, dt * f0, c3, c2, c1]


def dense_eval(coefs, t0, t1, t):
    th = (t - t0) / (t1 - t0)
    out = coefs[0] + th * coefs[1]
    p = th
    for cf in coefs[2:]:
        p = p * th
        out = out + p * cf   # reference uses in-place +=; same bits
    return out


# --------------------------------------------------------------------------------------
# Integration loops.  Reference: torchdyn/numerics/odeint.py:29-91 (dispatch),
# :326-408 (adaptive), :411-445 (fixed).
# --------------------------------------------------------------------------------------

class Trace:
    """Per-attempt record of one adaptive solve (not in the reference; used to pin step sequences)."""
    def __init__(self):
        self.t: List[float] = []
        self.dt: List[float] = []
        self.ratio: List[float] = []
        self.accept: List[bool] = []
        self.dt0: Optional[float] = None


def _adaptive(f, k1, x, dt, t_span, tab: Tableau, atol, rtol, interp_bmid, return_all_eval, seminorm,
              trace: Optional[Trace] = None, max_steps: int = 1_000_000):
    t_eval, t, T = t_span[1:], t_span[:1], t_span[-1]
    ck, ck_flag, dt_keep = 0, False, None
    times, states = [t], [x]
    n = 0
    while t < T:
        n += 1
        if n > max_steps:
            raise RuntimeError("oracle: max_steps exceeded")
        if t + dt > T:
            dt = T - t
        if ck < len(t_eval) and t + dt > t_eval[ck] and interp_bmid is None:
            dt_keep, ck_flag = dt, True          # odeint.py:357-361
            dt = t_eval[ck] - t
        f_new, x_new, err, ks = rk_step(tab, f, x, t, dt, k1)
        if seminorm[0]:
            m = seminorm[1]
            es = err[:m] / (atol + rtol * torch.max(x[:m].abs(), x_new[:m].abs()))
        else:
            es = err / (atol + rtol * torch.max(x.abs(), x_new.abs()))
        ratio = hairer_norm(es)
        ok = bool(ratio <= 1)
        if trace is not None:
            trace.t.append(float(t)); trace.dt.append(float(dt)); trace.ratio.append(float(ratio)); trace.accept.append(ok)
        if ok:
            if interp_bmid is not None:
                coefs = None
                while ck < len(t_eval) and t + dt > t_eval[ck]:
                    if coefs is None:
                        x_mid = x + dt * sum([interp_bmid[i] * ks[i] for i in range(7)])   # odeint.py:381
                        coefs = dense_fit_4th(dt, k1, f_new, x, x_new, x_mid)
                    states.append(dense_eval(coefs, t, t + dt, t_eval[ck]))
                    times.append(t_eval[ck][None])
                    ck += 1
            if t + dt == t_eval[ck] or return_all_eval:      # exact float equality (odeint.py:389)
                states.append(x_new)
                times.append(t + dt)
                if t + dt == t_eval[ck]:
                    ck += 1
            t, x, k1 = t + dt, x_new, f_new
        if ck_flag:
            dt = dt_keep - dt                                   # the remainder (odeint.py:399-401)
            ck_flag = False
        dt = adapt_step(dt, ratio, tab.order)
    return torch.cat(times), torch.stack(states)


def _fixed(f, x, t_span, tab: Tableau, save_at=()):
    if len(save_at) == 0:
        save_at = t_span
    if not isinstance(save_at, Tensor):
        save_at = torch.tensor(save_at)
    assert all(torch.isclose(s, save_at).sum() == 1 for s in save_at)
    t, dt = t_span[0], t_span[1] - t_span[0]
    out = [x] if torch.isclose(t, save_at).sum() else []
    for step in range(1, len(t_span)):
        _, x, _, _ = rk_step(tab, f, x, t, dt, None)          # no FSAL carry (odeint.py:428)
        t = t + dt
        if torch.isclose(t, save_at).sum():
            out.append(x)
        if step < len(t_span) - 1:
            dt = t_span[step + 1] - t                            # uses the ACCUMULATED t (odeint.py:433)
    return save_at, torch.stack(out)


def odeint(f, x, t_span, solver: str, atol=1e-3, rtol=1e-3, interpolator=None, return_all_eval=False,
           save_at=(), seminorm=(False, None), tableau_dtype=None, trace: Optional[Trace] = None):
    """Restatement of torchdyn.numerics.odeint (odeint.py:29-91) for the four hot-path solvers.
    ``tableau_dtype``: dtype the tableau is BUILT in before being cast to x.dtype -- float32 when the
    solver came from ODEProblem/NeuralODE by name (problems.py:57), x.dtype for the functional API (odeint.py:63)."""
    if t_span[1] < t_span[0]:
        g = lambda tt, xx: -f(-tt, xx)
        t_span = -t_span
    else:
        g = f
    tab = Tableau(solver, tableau_dtype or x.dtype).cast(x)
    t_span = t_span.to(x.device)
    bmid = None
    if interpolator is not None and tab.name != "tsit5":       # tsit5 drops the interpolator (odeint.py:68-70)
        if interpolator != "4th":
            raise KeyError(interpolator)
        bmid = bmid_4th(x.dtype)
    if t_span.dim() > 1:
        raise NotImplementedError("Parallel routines not implemented yet, check experimental versions of `torchdyn`")
    if not tab.adaptive:
        return _fixed(g, x, t_span, tab, save_at)
    t0 = t_span[0]
    k1 = g(t0, x)
    dt = init_step(f, k1, x, t0, tab.order, atol, rtol)        # un-negated f on purpose
    if trace is not None:
        trace.dt0 = float(dt)
    return _adaptive(g, k1, x, dt, t_span, tab, atol, rtol, bmid, return_all_eval, seminorm, trace)


# --------------------------------------------------------------------------------------
# Natural cubic spline -- restatement of the PUBLISHED torchcde 0.2.x algorithm
# (torchcde.natural_cubic_coeffs / CubicSpline.evaluate; package not present in the reference tree,
# constraint ``torchcde ^0.2.3`` in /root/reference/pyproject.toml:15).  Call sites:
# torchdyn/numerics/sensitivity.py:126-127 (build), :132 (evaluate).  PARITY UNPINNED.
# --------------------------------------------------------------------------------------

def _thomas(rhs: Tensor, off: Tensor, diag: Tensor) -> Tensor:
    """Solve the symmetric tridiagonal system along the last dim (sequential over knots)."""
    L = rhs.size(-1)
    dprime = [diag[0]]
    bprime = [rhs[..., 0]]
    for i in range(1, L):
        w = off[i - 1] / dprime[i - 1]
        dprime.append(diag[i] - w * off[i - 1])
        bprime.append(rhs[..., i] - w * bprime[i - 1])
    out = [None] * L
    out[L - 1] = bprime[L - 1] / dprime[L - 1]
    for i in range(L - 2, -1, -1):
        out[i] = (bprime[i] - off[i] * out[i + 1]) / dprime[i]
    return torch.stack(out, dim=-1)


def natural_cubic_coeffs(x: Tensor, t: Optional[Tensor] = None) -> Tensor:
    """x: [..., L, C], t: [L] -> coeffs [..., L-1, 4C] = cat(a, b, two_c, three_d)."""
    L = x.size(-2)
    if t is None:
        t = torch.linspace(0, L - 1, L, dtype=x.dtype, device=x.device)
    xt = x.transpose(-1, -2)                                     # [..., C, L]
    if L < 2:
        raise ValueError("need at least two knots")
    if L == 2:
        a = xt[..., :1]
        b = (xt[..., 1:] - xt[..., :1]) / (t[1:] - t[:1])
        two_c = torch.zeros_like(a)
        three_d = torch.zeros_like(a)
    else:
        h = t[1:] - t[:-1]
        r = h.reciprocal()
        r2 = r ** 2
        three_dx = 3 * (xt[..., 1:] - xt[..., :-1])
        six_dx = 2 * three_dx
        dx_scaled = three_dx * r2
        diag = torch.empty(L, dtype=x.dtype, device=x.device)
        diag[:-1] = r
        diag[-1] = 0
        diag[1:] += r
        diag *= 2
        rhs = torch.empty_like(xt)
        rhs[..., :-1] = dx_scaled
        rhs[..., -1] = 0
        rhs[..., 1:] += dx_scaled
        kd = _thomas(rhs, r, diag)
        a = xt[..., :-1]
        b = kd[..., :-1]
        two_c = (six_dx * r - 4 * kd[..., :-1] - 2 * kd[..., 1:]) * r
        three_d = (-six_dx * r + 3 * (kd[..., :-1] + kd[..., 1:])) * r2
    parts = [p.transpose(-1, -2) for p in (a, b, two_c, three_d)]
    return torch.cat(parts, dim=-1)


class NaturalCubicSpline:
    def __init__(self, a, b, two_c, three_d, t):
        self.a, self.b, self.two_c, self.three_d, self.t = a, b, two_c, three_d, t

    @classmethod
    def from_coeffs(cls, coeffs: Tensor, t: Tensor):
        C = coeffs.size(-1) // 4
        return cls(coeffs[..., :C], coeffs[..., C:2 * C], coeffs[..., 2 * C:3 * C], coeffs[..., 3 * C:], t)

    def evaluate(self, t):
        t = torch.as_tensor(t, dtype=self.b.dtype, device=self.b.device)
        last = self.b.size(-2) - 1
        idx = torch.bucketize(t.detach(), self.t.detach()).sub(1).clamp(0, last)   # clamp-index extrapolation
        s = (t - self.t[idx]).unsqueeze(-1)
        inner = 0.5 * self.two_c[..., idx, :] + self.three_d[..., idx, :] * s / 3
        inner = self.b[..., idx, :] + inner * s
        return self.a[..., idx, :] + inner * s


# --------------------------------------------------------------------------------------
# Sensitivities.  Reference: torchdyn/numerics/sensitivity.py:32-100 (backsolve adjoint),
# :104-161 (interpolated adjoint); parameter flattening torchdyn/core/problems.py:104-108.
# --------------------------------------------------------------------------------------

def _flat_params(vf) -> Tensor:
    return torch.cat([p.contiguous().flatten() for p in vf.parameters()])


def _vjp_pack(vf, t, x, lam, with_dt: bool, integral_loss=None):
    """(dx, dlam, dmu[, dt]) = (f, -lam^T df/dx [- dg/dx], -lam^T df/dtheta[, -lam^T df/dt])  (sensitivity.py:62-74;
    the running cost g = integral_loss(t, x) enters the co-state dynamics, :67-69 and :141-143)."""
    with torch.set_grad_enabled(True):
        x = x.requires_grad_(True)
        t = t.requires_grad_(True)
        dx = vf(t, x)
        params = tuple(vf.parameters())
        g = torch.autograd.grad(dx, (x, t) + params, -lam, allow_unused=True, retain_graph=False)
        dlam, dtt, dmu = g[0], g[1], g[2:]
        if integral_loss:
            dg = torch.autograd.grad(integral_loss(t, x).sum(), x, allow_unused=True, retain_graph=True)[0]
            dlam = dlam - dg
        dmu = torch.cat([e.flatten() if e is not None else torch.zeros(1, device=x.device) for e in dmu], dim=-1)
        if dtt is None:
            dtt = torch.zeros(1).to(t)
        if dtt.dim() == 0:
            dtt = dtt.unsqueeze(0)
    return dx, dlam, dmu, dtt


def make_adjoint_function(vf, solver, atol, rtol, interpolator, solver_adjoint, atol_adjoint, rtol_adjoint,
                          kind: str, traces: Optional[list] = None, integral_loss=None):
    """Build the autograd.Function for ``kind`` in {'adjoint', 'interpolated_adjoint'}.
    ``traces``: optional list; a Trace is appended per odeint call (forward first, then each backward interval)."""
    assert kind in ("adjoint", "interpolated_adjoint")

    def _tr():
        if traces is None:
            return None
        traces.append(Trace())
        return traces[-1]

    class _Backsolve(torch.autograd.Function):
        @staticmethod
        def forward(ctx, flat_params, x, t_span):
            t_sol, sol = odeint(vf, x, t_span, solver, atol, rtol, interpolator, False,
                                tableau_dtype=torch.float32, trace=_tr())
            ctx.save_for_backward(sol, t_sol)
            return t_sol, sol

        @staticmethod
        def backward(ctx, g_t, g_sol):
            sol, t_sol = ctx.saved_tensors
            xT, lamT = sol[-1], g_sol[-1]
            P = _flat_params(vf).numel()
            n = xT.numel()
            lam_flat = lamT.flatten()
            lam_t = lam_flat @ vf(t_sol[-1], xT).flatten()                   # sensitivity.py:53
            A = torch.cat([xT.flatten(), lam_flat, torch.zeros(P, dtype=xT.dtype, device=xT.device), lam_t[None]])

            def aug(t, A):                                                     # sensitivity.py:57-75
                if t.dim() > 0:
                    t = t[0]
                x = A[:n].reshape(xT.shape)
                lam = A[n:2 * n].reshape(lamT.shape)
                dx, dlam, dmu, dtt = _vjp_pack(vf, t, x, lam, True, integral_loss)
                return torch.cat([dx.flatten(), dlam.flatten(), dmu.flatten(), dtt.flatten()])

            dLdt = torch.zeros(len(t_sol)).to(xT)
            dLdt[-1] = lam_t
            for i in range(len(t_sol) - 1, 0, -1):
                _, As = odeint(aug, A, t_sol[i - 1:i + 1].flip(0), solver_adjoint, atol_adjoint, rtol_adjoint,
                               seminorm=(True, 2 * n), tableau_dtype=torch.float32, trace=_tr())
                last = As[-1]
                xt = last[:n].reshape(xT.shape)
                dLdt[i - 1] = last[n:2 * n] @ vf(t_sol[i], xt).flatten()      # index i, not i-1 (sensitivity.py:88)
                last[-1:] -= g_t[i - 1]
                A = torch.cat([last[:n], last[n:2 * n], last[-P - 1:-1], last[-1:]])
                A[n:2 * n] += g_sol[i - 1].flatten()
            lam = A[n:2 * n].reshape(lamT.shape)
            mu = A[-P - 1:-1]
            return mu, lam, torch.stack([dLdt[0], dLdt[-1]])

    class _Interp(torch.autograd.Function):
        @staticmethod
        def forward(ctx, flat_params, x, t_span):
            t_sol, sol = odeint(vf, x, t_span, solver, atol, rtol, interpolator, True,
                                tableau_dtype=torch.float32, trace=_tr())
            ctx.save_for_backward(sol, t_span, t_sol)
            return t_sol, sol

        @staticmethod
        def backward(ctx, g_t, g_sol):
            sol, t_span, t_sol = ctx.saved_tensors
            xT, lamT = sol[-1], g_sol[-1]
            P = _flat_params(vf).numel()
            n = lamT.numel()
            A = torch.cat([lamT.flatten(), torch.zeros(P, dtype=xT.dtype, device=xT.device)])
            spline = NaturalCubicSpline.from_coeffs(natural_cubic_coeffs(sol.permute(1, 0, 2).detach(), t_sol), t_sol)

            def aug(t, A):                                                     # sensitivity.py:130-147
                if t.dim() > 0:
                    t = t[0]
                x = spline.evaluate(t)
                lam = A[:n].reshape(lamT.shape)
                _, dlam, dmu, _ = _vjp_pack(vf, t, x, lam, False, integral_loss)
                return torch.cat([dlam.flatten(), dmu.flatten()])

            for i in range(len(t_span) - 1, 0, -1):
                # forward solver / tolerances, no seminorm (sensitivity.py:152)
                _, As = odeint(aug, A, t_span[i - 1:i + 1].flip(0), solver, atol, rtol,
                               tableau_dtype=torch.float32, trace=_tr())
                last = As[-1]
                A = torch.cat([last[:n], last[-P:]])
                A[:n] += g_sol[i - 1].flatten()
            return A[-P:], A[:n].reshape(lamT.shape), None

    return _Backsolve if kind == "adjoint" else _Interp


class OracleNeuralODE(torch.nn.Module):
    """Minimal CPU stand-in for NeuralODE(vf, solver, sensitivity=...) forward(x, t_span) -> (t_eval, sol)
    (core/neuralde.py:29-136, core/problems.py:104-159) used as checker and CPU baseline.
    ``vf`` is called as vf(t, x); modules taking only x are wrapped; ``nfe`` counts calls (defunc.py:61)."""

    def __init__(self, vector_field, solver="tsit5", atol=1e-3, rtol=1e-3, sensitivity="autograd",
                 solver_adjoint=None, atol_adjoint=1e-4, rtol_adjoint=1e-4, interpolator=None, integral_loss=None):
        super().__init__()
        import inspect
        self.integral_loss = integral_loss
        self.net = vector_field
        fwd = vector_field.forward if isinstance(vector_field, torch.nn.Module) else vector_field
        self._has_t = "t" in inspect.getfullargspec(fwd).args
        self.nfe = 0
        self.solver, self.atol, self.rtol = solver, atol, rtol
        self.solver_adjoint = solver_adjoint or solver
        self.atol_adjoint, self.rtol_adjoint = atol_adjoint, rtol_adjoint
        self.sensitivity, self.interpolator = sensitivity, interpolator
        self.traces: list = []

    def parameters(self, recurse=True):
        return self.net.parameters() if isinstance(self.net, torch.nn.Module) else iter(())

    def vf(self, t, x):
        self.nfe += 1
        if self.integral_loss is not None and self.sensitivity == "autograd":
            # the running cost is integrated in column 0 of the (user-augmented) state (core/defunc.py:68-77)
            x_dyn = x[:, 1:]
            dlds = self.integral_loss(t, x_dyn)
            if len(dlds.shape) == 1:
                dlds = dlds[:, None]
            f = self.net(t, x_dyn) if self._has_t else self.net(x_dyn)
            return torch.cat([dlds, f], 1).to(f)
        return self.net(t, x) if self._has_t else self.net(x)

    def forward(self, x, t_span=None):
        if t_span is None:
            t_span = torch.linspace(0, 1, 2)
        self.traces = []
        vfw = _VF(self)
        if self.sensitivity == "autograd":
            tr = Trace(); self.traces.append(tr)
            return odeint(vfw, x, t_span, self.solver, self.atol, self.rtol, self.interpolator,
                          tableau_dtype=torch.float32, trace=tr)
        fn = make_adjoint_function(vfw, self.solver, self.atol, self.rtol, self.interpolator, self.solver_adjoint,
                                   self.atol_adjoint, self.rtol_adjoint, self.sensitivity, self.traces, self.integral_loss)
        return fn.apply(_flat_params(vfw), x, t_span)


class _VF:
    def __init__(self, owner: OracleNeuralODE):
        self.o = owner

    def __call__(self, t, x):
        return self.o.vf(t, x)

    def parameters(self):
        return self.o.parameters()
