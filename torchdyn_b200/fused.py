"""Recognition of fusable MLP vector fields and dispatch to the fused kernels.

A vector field is *fusable* when, after unwrapping torchdyn's call-signature adapters (``DEFunc`` ->
``DEFuncBase(has_time_arg=False)``, reference core/defunc.py:19-94) or this module's ``MLPField``, it is an
``nn.Sequential`` of ``Linear`` layers separated by one kind of activation out of Tanh / Softplus(beta=1,
threshold=20) / ReLU, time-independent, fp32 on CUDA.  The fused kernels evaluate the field inside this library
(``tdyn_mlp_tc_stage``: tcgen05, split precision TF32 + bf16 corrections), so ``f`` is never called; the wrappers' ``nfe`` counters are advanced by
the same amounts the reference would count (core/defunc.py:31,61).
"""
from __future__ import annotations

import weakref
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _kernels, _lib
from ._lib import ACT_RELU, ACT_SOFTPLUS, ACT_TANH, MODE_INNER

f32 = np.float32
DISABLED = False          # bench.py --force-generic / A-B parity runs: keep f a torch call
_last = None


def last_used():
    """Name of the fused path the most recent solve took (None = generic kernels, f via torch)."""
    return _last


def reset_last():
    global _last
    _last = None



def _bump(counters, n):
    """``c.nfe += n`` for every wrapper in ``counters``.  ``nfe`` is a plain float attribute; writing it through
    ``nn.Module.__setattr__`` costs ~6 us per wrapper and call, so it goes straight into the instance dict."""
    for c in counters:
        d = c.__dict__
        if "nfe" in d:
            d["nfe"] += n
        else:
            c.nfe += n


class MLPField(nn.Module):
    """Functional-API helper: ``odeint(MLPField(net), x, t_span, ...)`` == ``odeint(lambda t, x: net(x), ...)`` but
    lets the library see (and fuse) the network."""

    def __init__(self, net: nn.Sequential):
        super().__init__()
        self.net = net
        self.nfe = 0.0

    def forward(self, t, x, args={}):
        self.nfe += 1
        return self.net(x)


def _unwrap(f):
    """-> (nn.Sequential | None, [objects whose .nfe must advance])"""
    from .core.defunc import DEFunc, DEFuncBase
    counters = []
    cur = f
    for _ in range(4):
        if isinstance(cur, MLPField):
            counters.append(cur)
            return cur.net, counters
        if isinstance(cur, DEFunc):
            if cur.order > 1 or (cur.integral_loss is not None and cur.sensitivity == "autograd"):
                return None, []
            counters.append(cur)
            cur = cur.vf
            continue
        if isinstance(cur, DEFuncBase):
            if cur.has_time_arg:
                return None, []
            counters.append(cur)
            cur = cur.vf
            continue
        break
    return (cur, counters) if isinstance(cur, nn.Sequential) and counters else (None, [])


def _unwrap_cnf(f):
    """-> (CNF module | None, [objects whose .nfe must advance]) for DEFunc -> DEFuncBase -> CNF(net, hutch_trace)."""
    from .core.defunc import DEFunc, DEFuncBase
    from .models.cnf import CNF, hutch_trace
    counters = []
    cur = f
    for _ in range(3):
        if isinstance(cur, DEFunc):
            if cur.order > 1 or (cur.integral_loss is not None and cur.sensitivity == "autograd"):
                return None, []
            counters.append(cur)
            cur = cur.vf
            continue
        if isinstance(cur, DEFuncBase):
            if cur.has_time_arg:
                return None, []
            counters.append(cur)
            cur = cur.vf
            continue
        break
    if isinstance(cur, CNF) and counters and cur.trace_estimator is hutch_trace and cur.order == 1 \
            and isinstance(cur.net, nn.Sequential):
        return cur, counters
    return None, []


_cache = weakref.WeakKeyDictionary()      # nn.Sequential -> (ids of its modules, FusedMLP | False); kept off the module: it must stay picklable
_ACTS = {nn.Tanh: ACT_TANH, nn.Softplus: ACT_SOFTPLUS, nn.ReLU: ACT_RELU}


class FusedMLP:
    """Descriptor (``tdyn_mlp_t``) + prepared weight images of one fusable network, cached on the module."""

    def __init__(self, net: nn.Sequential):
        mods = list(net)
        if len(mods) < 3 or len(mods) % 2 == 0:
            raise ValueError("not Linear(-act-Linear)+")
        linears, acts = mods[0::2], mods[1::2]
        if not all(type(m) is nn.Linear and m.bias is not None for m in linears):
            raise ValueError("non-Linear layer")
        kinds = {type(a) for a in acts}
        if len(kinds) != 1 or next(iter(kinds)) not in _ACTS:
            raise ValueError("unsupported activation")
        a0 = acts[0]
        if isinstance(a0, nn.Softplus) and not all(a.beta == 1 and a.threshold == 20 for a in acts):
            raise ValueError("Softplus must have beta=1, threshold=20")
        if len(linears) > _lib.MAX_LAYERS:
            raise ValueError("too deep")
        for a, b in zip(linears[:-1], linears[1:]):
            if a.out_features != b.in_features:
                raise ValueError("shape chain broken")
        if linears[0].in_features != linears[-1].out_features:
            raise ValueError("not an autonomous vector field")
        self.linears = linears
        self.act = _ACTS[type(a0)]
        self.dims = [linears[0].in_features] + [m.out_features for m in linears]
        self.flops_per_row = 2.0 * sum(a * b for a, b in zip(self.dims[:-1], self.dims[1:]))
        self.n_params = sum(a * b + b for a, b in zip(self.dims[:-1], self.dims[1:]))
        self._key = None
        self._desc = None
        self._img = None
        self._tc_ok = False
        self._ffma_ok = False
        self._ffma_ws = None
        self._wide_ok = False
        self._wide_imgs = None

    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for m in self.linears for p in (m.weight, m.bias))

    def begin_solve(self):
        """Called once per user-level solve / backward: forget the prepared weight images, so that they are re-packed
        from the live parameters.  ``(data_ptr, _version)`` alone misses in-place updates through ``.data`` (weight
        clipping, EMA, older optimizers), and re-packing costs a few microseconds next to a solve."""
        self._key = None
        self._fresh = False

    def _plain(self) -> bool:
        """Only untouched nn.Linear layers are fused: a forward / pre-forward hook (old-style weight_norm / spectral_norm
        recompute ``weight`` there) or a non-Parameter weight means torch must evaluate the layer."""
        for m in self.linears:
            if m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or m._backward_pre_hooks:
                return False
            if not (isinstance(m.weight, nn.Parameter) and isinstance(m.bias, nn.Parameter)):
                return False
        return True

    def ready(self, kern) -> bool:
        """True if the tensor-core stage kernel can run this net (weight image rebuilt if the parameters changed)."""
        return self._refresh(kern) and self._tc_ok

    def ready_ffma(self, kern) -> bool:
        """True if the narrow-net FFMA kernels can run this net."""
        return self._refresh(kern) and self._ffma_ok

    def ready_wide(self, kern) -> bool:
        """True if the layer-wise tensor-core kernels (tdyn_lin_tc_apply / tdyn_wgrad_tc) can run this net."""
        return self._refresh(kern) and self._wide_ok

    def wide_images(self, kern):
        """[(forward image, input-gradient image)] per layer, rebuilt after a parameter update."""
        if self._wide_imgs is None:
            self._wide_imgs = [(kern.lin_tc_prepare(m.weight, False), kern.lin_tc_prepare(m.weight, True)) for m in self.linears]
        return self._wide_imgs

    def tc_transposed(self, kern):
        """(descriptor, prepared image) of the TRANSPOSED 32-256-256-32 network: the input-gradient chain of the fused
        stage kernel (``tdyn_mlp_tc_dgrad``); rebuilt after a parameter update."""
        if self.__dict__.get("_tc_t") is None:
            wts = [m.weight.detach().t().contiguous() for m in reversed(self.linears)]      # W3^T, W2^T, W1^T
            zeros = [torch.zeros(w.shape[0], dtype=torch.float32, device=w.device) for w in wts]
            d = _lib.MlpDesc()
            d.n_layers = len(wts)
            for i, v in enumerate(reversed(self.dims)):
                d.dims[i] = v
            d.act = self.act
            for i, (w, b) in enumerate(zip(wts, zeros)):
                d.w[i] = w.data_ptr()
                d.b[i] = b.data_ptr()
            self._tc_t = (d, kern.mlp_tc_prepare(d), wts, zeros)
        return self._tc_t[0], self._tc_t[1]

    def scratch(self, key, make):
        """Per-network cache of solver scratch buffers (stage storage, ping-pong state, status, trace): they are
        consumed within the call, so they can be reused by the next one (stream-ordered)."""
        cache = self.__dict__.setdefault("_scratch", {})
        buf = cache.get(key)
        if buf is None:
            if len(cache) > 8:
                cache.clear()
            buf = cache[key] = make()
        return buf

    def solve_workspace(self, kern):
        if getattr(self, "_solve_ws", None) is None or self._solve_ws.device != kern.device:
            self._solve_ws = kern.mlp_solve_workspace(self._desc)
        return self._solve_ws

    def workspace(self, kern):
        if self._ffma_ws is None or self._ffma_ws.device != kern.device:
            self._ffma_ws = kern.mlp_ffma_workspace(self._desc)
        return self._ffma_ws

    def _refresh(self, kern) -> bool:
        # within one solve (between two begin_solve calls) the first check stands: this runs once per f evaluation
        if self.__dict__.get("_fresh") and self._fresh_dev == kern.device:
            return self._fresh_ok
        ok = self._refresh_now(kern)
        self._fresh, self._fresh_dev, self._fresh_ok = True, kern.device, ok
        return ok

    def _refresh_now(self, kern) -> bool:
        if not self._plain():
            return False
        ps = [p for m in self.linears for p in (m.weight, m.bias)]
        if not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in ps):
            return False
        key = self._params_key()
        if key == self._key:
            return True
        d = _lib.MlpDesc()
        d.n_layers = len(self.linears)
        for i, v in enumerate(self.dims):
            d.dims[i] = v
        d.act = self.act
        for i, m in enumerate(self.linears):
            d.w[i] = m.weight.data_ptr()
            d.b[i] = m.bias.data_ptr()
        self._key, self._desc, self._img, self._wide_imgs, self._tc_t = key, d, None, None, None
        if self.__dict__.get("_shape_flags") is None:
            # which kernel family takes this net depends on its layer widths and activation only
            ffma = kern.mlp_ffma_supported(d)
            self._shape_flags = (kern.mlp_tc_supported(d), ffma,
                                 not ffma and all(kern.lin_tc_supported(o, i) and kern.lin_tc_supported(i, o)
                                                  for i, o in zip(self.dims[:-1], self.dims[1:])))
        self._tc_ok, self._ffma_ok, self._wide_ok = self._shape_flags
        if self._tc_ok:
            self._img = kern.mlp_tc_prepare(d)
        return True


def _fused_of(net: nn.Sequential):
    """FusedMLP of ``net`` (or False), re-derived when the Sequential's module list was mutated after first use."""
    ids = tuple(id(m) for m in net)
    hit = _cache.get(net)
    if hit is not None and hit[0] == ids:
        return hit[1]
    try:
        fm = FusedMLP(net)
    except ValueError:
        fm = False
    _cache[net] = (ids, fm)
    return fm


def _get_fused(f) -> Optional[tuple]:
    net, counters = _unwrap(f)
    if net is None:
        return None
    fm = _fused_of(net)
    return (fm, counters) if fm else None


class FusedField:
    """``f`` with a kernel-backed ``tdyn_eval``: the integration loops evaluate the MLP inside this library (FFMA
    kernel for narrow nets, the tensor-core stage kernel for the 32-256-256-32 class, the layer-wise tensor-core
    kernels for other wide nets) instead of calling torch.
    ``__call__`` still forwards to the original callable (user code may call it directly)."""

    def __init__(self, f, fm: FusedMLP, counters, kern):
        self.f, self.fm, self.counters, self.kern = f, fm, counters, kern

    def __call__(self, t, x, *a, **k):
        return self.f(t, x, *a, **k)

    def parameters(self):
        return self.f.parameters()

    def tdyn_eval(self, t_host, y, sign: float = 1.0):
        fm, kern = self.fm, self.kern
        if y.dim() != 2 or y.shape[1] != fm.dims[0] or not y.is_contiguous():
            out = self.f(kern.time_tensor(t_host, y), y)
            return out if sign == 1.0 else -out
        out = torch.empty_like(y)
        if fm.ready(kern) and sign == 1.0:
            kern.mlp_tc_stage(fm._desc, fm._img, out, y, [], (), MODE_INNER, 0.0, None, (), fm.flops_per_row)
        elif fm.ready_ffma(kern):
            kern.mlp_ffma_forward(fm._desc, y, out, sign, fm.flops_per_row)
        else:
            wide_forward(fm, kern, y, out, sign)
        _bump(self.counters, 1)
        return out


LIN_MODE_LINEAR, LIN_MODE_ACT, LIN_MODE_DACT = 0, 1, 2


def _aligned(t):
    """The tensor-core kernels use 128-bit accesses: views at odd offsets of a flat vector are copied."""
    return t if t.data_ptr() % 16 == 0 else t.clone()


def wide_forward(fm: "FusedMLP", kern, y, out, sign: float = 1.0):
    """Layer-by-layer forward of a wide MLP on the tensor cores (``tdyn_lin_tc_apply``, TF32 + bf16 corrections): hidden layers write
    their activation outputs (returned: the adjoint reuses them), the last layer writes ``sign * f`` into ``out``
    (skipped when ``out`` is None: the interpolated adjoint does not need f itself)."""
    imgs, dims, rows = fm.wide_images(kern), fm.dims, y.numel() // fm.dims[0]
    L = len(fm.linears)
    if fm.ready(kern) and rows > 0:
        # the 32-256-256-32 class: ONE launch of the fused stage kernel, hidden activations written out on the way
        # (also when the caller does not need f itself: the same kernel, so the same activations to the last bit)
        ya = _aligned(y)
        o = out if out is not None and out.data_ptr() % 16 == 0 else torch.empty(rows, dims[-1], dtype=torch.float32, device=y.device)
        a1 = torch.empty(rows, dims[1], dtype=torch.float32, device=y.device)
        a2 = torch.empty(rows, dims[2], dtype=torch.float32, device=y.device)
        kern.mlp_tc_forward_save(fm._desc, fm._img, o, ya, sign, a1, a2, fm.flops_per_row)
        if out is not None and o is not out:
            out.view(-1).copy_(o.view(-1))
        return [a1, a2]
    h, saved = _aligned(y), []
    for l, lin in enumerate(fm.linears):
        last = l == L - 1
        if last and out is None:
            break
        o = torch.empty(rows, dims[l + 1], dtype=torch.float32, device=y.device) if not last or out.data_ptr() % 16 else out
        kern.lin_tc_apply(imgs[l][0], dims[l + 1], dims[l], h, lin.bias, None, o, fm.act,
                          LIN_MODE_LINEAR if last else LIN_MODE_ACT, sign if last else 1.0, rows)
        if not last:
            saved.append(o)
        elif o is not out:
            out.view(-1).copy_(o.view(-1))
        h = o
    return saved


class FusedWideAdjoint:
    """Body of the adjoint dynamics for a wide MLP on the tensor cores: forward (activations kept), then per layer,
    last to first, the parameter gradient ``-lam^T df/dtheta`` contracted over the batch (``tdyn_wgrad_tc``) and the
    input gradient (``tdyn_lin_tc_apply`` with the transposed weight image, times act' of the saved activations).
    Same call signature as FusedAdjoint; ``dx_out`` may be None (interpolated adjoint)."""
    f_optional = True

    def __init__(self, fm: "FusedMLP", counters, kern):
        self.fm, self.counters, self.kern = fm, counters, kern

    def __call__(self, x_flat, lam_flat, dx_out, dlam_out, dmu_out, rows, sign):
        fm, kern = self.fm, self.kern
        dims, imgs = fm.dims, fm.wide_images(kern)
        x_flat, lam_flat = _aligned(x_flat), _aligned(lam_flat)
        saved = wide_forward(fm, kern, x_flat, dx_out, sign)
        offs, o_ = [], 0
        for i, o in zip(dims[:-1], dims[1:]):
            offs.append(o_)
            o_ += o * i + o
        if fm.ready(kern) and len(saved) == 2:
            # the 32-256-256-32 class: the whole input-gradient chain in ONE launch of the fused stage kernel on the
            # transposed network; its hidden cotangents d2, d1 and lam feed the three parameter-gradient contractions
            a1, a2 = saved
            d2, d1 = torch.empty_like(a2), torch.empty_like(a1)
            dl = dlam_out if dlam_out.data_ptr() % 16 == 0 else torch.empty(rows, dims[0], dtype=torch.float32, device=a1.device)
            desc_t, img_t = fm.tc_transposed(kern)
            kern.mlp_tc_dgrad(desc_t, img_t, dl, lam_flat, -sign, a2, a1, d2, d1, fm.flops_per_row)
            if dl is not dlam_out:
                dlam_out.view(-1).copy_(dl.view(-1))
            for l, (gl, inp) in enumerate(((d1, x_flat), (d2, a1), (lam_flat, a2))):
                i, o = dims[l], dims[l + 1]
                ws = fm.scratch(("wgrad", l, rows, x_flat.device.index), lambda: kern.wgrad_tc_workspace(o, i, rows))
                kern.wgrad_tc(gl, o, inp, i, dmu_out[offs[l]:offs[l] + o * i], dmu_out[offs[l] + o * i:offs[l] + o * i + o], -sign, ws, rows)
            _bump(self.counters, 1)
            return
        g = lam_flat
        for l in range(len(fm.linears) - 1, -1, -1):
            i, o = dims[l], dims[l + 1]
            inp = saved[l - 1] if l > 0 else x_flat
            ws = fm.scratch(("wgrad", l, rows, x_flat.device.index), lambda: kern.wgrad_tc_workspace(o, i, rows))
            kern.wgrad_tc(g, o, inp, i, dmu_out[offs[l]:offs[l] + o * i], dmu_out[offs[l] + o * i:offs[l] + o * i + o], -sign, ws, rows)
            if l > 0:
                gp = torch.empty(rows, i, dtype=torch.float32, device=x_flat.device)
                kern.lin_tc_apply(imgs[l][1], i, o, g, None, saved[l - 1], gp, fm.act, LIN_MODE_DACT, 1.0, rows)
                g = gp
            else:
                dl = dlam_out if dlam_out.data_ptr() % 16 == 0 else torch.empty(rows, i, dtype=torch.float32, device=g.device)
                kern.lin_tc_apply(imgs[0][1], i, o, g, None, None, dl, fm.act, LIN_MODE_LINEAR, -sign, rows)
                if dl is not dlam_out:
                    dlam_out.view(-1).copy_(dl.view(-1))
        _bump(self.counters, 1)


class FusedCNFField:
    """CNF(net, hutch_trace) evaluated by ``tdyn_cnf_ffma_forward``: MLP forward + the eps^T J VJP + the trace dot
    product in one launch (the reference builds an autograd graph per evaluation)."""

    def __init__(self, f, cnf, fm: FusedMLP, counters, kern):
        self.f, self.cnf, self.fm, self.counters, self.kern = f, cnf, fm, counters, kern

    def __call__(self, t, x, *a, **k):
        return self.f(t, x, *a, **k)

    def parameters(self):
        return self.f.parameters()

    def tdyn_eval(self, t_host, s, sign: float = 1.0):
        fm, kern, eps = self.fm, self.kern, self.cnf.noise
        D = fm.dims[0]
        if not (s.dim() == 2 and s.shape[1] == D + 1 and s.is_contiguous() and isinstance(eps, torch.Tensor) and eps.is_cuda
                and eps.dtype == torch.float32 and tuple(eps.shape) == (s.shape[0], D) and eps.is_contiguous()
                and fm.ready_ffma(kern)):
            out = self.f(kern.time_tensor(t_host, s), s)
            return out if sign == 1.0 else -out
        out = kern.cnf_ffma_forward(fm._desc, s, eps, torch.empty_like(s), sign, fm.flops_per_row)
        _bump(self.counters, 1)
        return out


class FusedCNFAdjoint:
    """Body of the adjoint dynamics of a CNF (forward, VJP, and the second-order terms) in one launch
    (``tdyn_cnf_ffma_adjoint``); same call signature as FusedAdjoint."""

    def __init__(self, cnf, fm: FusedMLP, counters, kern):
        self.cnf, self.fm, self.counters, self.kern = cnf, fm, counters, kern

    def __call__(self, s_flat, lam_flat, ds_out, dlam_out, dmu_out, rows, sign):
        fm = self.fm
        self.kern.cnf_ffma_adjoint(fm._desc, s_flat, lam_flat, self.cnf.noise, ds_out, dlam_out, dmu_out, fm.workspace(self.kern),
                                   rows, sign, fm.flops_per_row)
        _bump(self.counters, 1)


def _get_fused_cnf(f):
    cnf, counters = _unwrap_cnf(f)
    if cnf is None:
        return None
    fm = _fused_of(cnf.net)
    return (cnf, fm, counters) if fm else None


def wrap(f, x):
    """Return a FusedField for ``f`` when it is a fusable MLP on this device, else ``f`` unchanged."""
    global _last
    if DISABLED or isinstance(f, (FusedField, FusedCNFField)) or not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
        return f
    gc = _get_fused_cnf(f)
    if gc is not None:
        cnf, fm, counters = gc
        fm.begin_solve()
        kern = _kernels.get(x)
        if getattr(kern, "is_cuda", False) and x.dim() == 2 and x.shape[1] == fm.dims[0] + 1 and fm.ready_ffma(kern) \
                and kern.cnf_ffma_supported(fm._desc):
            _last = "fused CNF / Hutchinson vector field (tdyn_cnf_ffma_forward)"
            return FusedCNFField(f, cnf, fm, counters, kern)
        return f
    got = _get_fused(f)
    if got is None:
        return f
    fm, counters = got
    fm.begin_solve()
    kern = _kernels.get(x)
    if not getattr(kern, "is_cuda", False) or x.dim() != 2 or x.shape[1] != fm.dims[0]:
        return f
    if not (fm.ready(kern) or fm.ready_ffma(kern) or fm.ready_wide(kern)):
        return f
    _last = ("fused MLP vector field (tdyn_mlp_tc_stage)" if fm.ready(kern) else
             "fused MLP vector field (tdyn_mlp_ffma_forward)" if fm.ready_ffma(kern) else
             "wide MLP vector field, layer-wise tcgen05 split-precision kernels (tdyn_lin_tc_apply)")
    return FusedField(f, fm, counters, kern)


def eval_field(vf, t, x):
    """``vf(t, x)`` for the adjoint's boundary terms (sensitivity.py:53, :88): the FFMA kernel when ``vf`` is a
    fusable MLP (one launch, NFE counted as the wrappers would), the module itself otherwise."""
    if not DISABLED and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 \
            and x.is_contiguous():
        got = _get_fused(vf)
        if got is not None:
            fm, counters = got
            kern = _kernels.get(x)
            if getattr(kern, "is_cuda", False) and x.shape[1] == fm.dims[0] and (fm.ready_ffma(kern) or fm.ready_wide(kern)):
                out = torch.empty_like(x)
                if fm.ready_ffma(kern):
                    kern.mlp_ffma_forward(fm._desc, x, out, 1.0, fm.flops_per_row)
                else:
                    wide_forward(fm, kern, x, out, 1.0)
                _bump(counters, 1)
                return out
    return vf(t, x)


class FusedAdjoint:
    """Kernel-backed body of the adjoint dynamics for a fusable MLP: forward + VJP + batch-reduced parameter gradient
    in one launch (``tdyn_mlp_ffma_adjoint``)."""

    def __init__(self, fm: FusedMLP, counters, kern):
        self.fm, self.counters, self.kern = fm, counters, kern

    def __call__(self, x_flat, lam_flat, dx_out, dlam_out, dmu_out, rows, sign):
        fm = self.fm
        self.kern.mlp_ffma_adjoint(fm._desc, x_flat, lam_flat, dx_out, dlam_out, dmu_out, fm.workspace(self.kern), rows,
                                   sign, fm.flops_per_row)
        _bump(self.counters, 1)


def adjoint_for(vf, x_shape, device) -> Optional[FusedAdjoint]:
    """FusedAdjoint for ``vf`` (as held by ODEProblem: DEFunc / DEFuncBase wrapped nn.Sequential) or None."""
    if DISABLED or len(x_shape) != 2 or device.type != "cuda":
        return None
    gc = _get_fused_cnf(vf)
    if gc is not None:
        cnf, fm, counters = gc
        fm.begin_solve()
        kern = _kernels.get(torch.empty(0, device=device))
        eps = cnf.noise
        ok = (getattr(kern, "is_cuda", False) and x_shape[1] == fm.dims[0] + 1 and fm.ready_ffma(kern)
              and kern.cnf_ffma_supported(fm._desc) and isinstance(eps, torch.Tensor) and eps.is_cuda
              and eps.dtype == torch.float32 and eps.is_contiguous() and tuple(eps.shape) == (x_shape[0], fm.dims[0])
              and sum(p.numel() for p in vf.parameters()) == fm.n_params)
        return FusedCNFAdjoint(cnf, fm, counters, kern) if ok else None
    got = _get_fused(vf)
    if got is None:
        return None
    fm, counters = got
    if x_shape[1] != fm.dims[0]:
        return None
    fm.begin_solve()
    kern = _kernels.get(torch.empty(0, device=device))
    if not getattr(kern, "is_cuda", False) or not (fm.ready_ffma(kern) or fm.ready_wide(kern)):
        return None
    # the flat parameter vector of the problem must be exactly this net's parameters, in order
    if sum(p.numel() for p in vf.parameters()) != fm.n_params:
        return None
    return FusedAdjoint(fm, counters, kern) if fm.ready_ffma(kern) else FusedWideAdjoint(fm, counters, kern)


# ------------------------------------------------------------------------------------------------------------------
# persistent solver: a whole adaptive odeint call in one cooperative kernel (tdyn_mlp_solve)
# ------------------------------------------------------------------------------------------------------------------
PERSISTENT = True         # set False to keep the host-driven loop (A/B runs)
SOLVE_KERNEL = True       # fixed-step solves of the 32-256-256-32 class: one launch per solve (False: one per stage)
TRACE_CAP = 3 * 4096 + 1
MAX_STEPS = 4000          # attempts one launch may take (its trace buffer holds 4096); longer solves go to the host loop
MAX_T_SPAN = 1024         # tdyn_mlp_solve keeps t_span in a 4 KB workspace slot


def _rk_plan(solver) -> "_lib.RkPlan":
    p = solver.plan()
    cached = solver.__dict__.get("_tdyn_rk_plan")
    if cached is not None and cached[0] is p:
        return cached[1]
    rp = _lib.RkPlan()
    rp.n_extra = len(p.stages)
    for s, st in enumerate(p.stages):
        rp.mode[s], rp.nk[s], rp.c[s] = st.mode, len(st.coefs), float(st.c)
        for j, v in enumerate(st.coefs):
            rp.a[s][j] = v
    rp.n_sol = len(p.bsol)
    for j, v in enumerate(p.bsol):
        rp.bsol[j] = v
    rp.n_err = 0 if p.berr is None else len(p.berr)
    for j, v in enumerate(p.berr or ()):
        rp.berr[j] = v
    rp.order = solver.order
    rp.safety, rp.min_factor, rp.max_factor = solver.controller_constants()
    solver.__dict__["_tdyn_rk_plan"] = (p, rp)
    return rp


def _emit_trace(trace_dev, n_attempt):
    from .numerics import odeint as _  # noqa: F401  (module object lives in sys.modules)
    import sys
    om = sys.modules["torchdyn_b200.numerics.odeint"]
    if om.TRACE_SINK is None:
        return
    tr = trace_dev[:3 * n_attempt + 1].cpu().tolist()
    rec = {"dt0": tr[0], "t": tr[1::3], "dt": tr[2::3], "ratio": tr[3::3]}
    rec["accept"] = [r <= 1 for r in rec["ratio"]]
    om.TRACE_SINK.append(rec)


def _solver_ok(solver) -> bool:
    from .numerics.solvers.ode import DormandPrince45, Tsitouras45
    return type(solver) in (DormandPrince45, Tsitouras45)


def try_adaptive(f_, x, ts, solver, atol, rtol, interpolator, return_all_eval, seminorm):
    """Adaptive solve of a fusable narrow MLP field in ONE kernel launch.  ``f_`` is what odeint integrates (a
    FusedField, or Reversed(FusedField) on a reversed span).  Returns (t_eval, sol) or None (generic path)."""
    global _last
    from .numerics.solvers.ode import Reversed
    from . import distributed as _dist
    if DISABLED or not PERSISTENT or interpolator is not None or seminorm[0] or not _solver_ok(solver):
        return None
    sign, inner = 1.0, f_
    if isinstance(f_, Reversed):
        sign, inner = -1.0, f_.f
    if not isinstance(inner, FusedField):
        return None
    fm, kern = inner.fm, inner.kern
    if not (x.dim() == 2 and x.is_contiguous() and x.shape[1] == fm.dims[0]):
        return None
    if fm.ready(kern):
        return _try_adaptive_tc(inner, sign, x, ts, solver, atol, rtol, return_all_eval)
    if not kern.mlp_solve_supported(fm._desc, False):
        return None
    comm = None
    if _dist.is_enabled():
        comm = _dist.peer_comm(x.device)      # in-kernel NVLink exchange of the error norm; None -> NCCL host path
        if comm is None:
            return None
    rows, D = x.shape
    n_t = len(ts)
    if n_t > MAX_T_SPAN:
        return None
    cap = n_t + (512 if return_all_eval else 0)
    dev = x.device
    y0 = x.clone()
    y1, k, status, trace = fm.scratch(("fwd", rows, dev.index), lambda: (
        torch.empty_like(x), torch.empty(7 * rows * D, dtype=torch.float32, device=dev),
        torch.zeros(8, dtype=torch.int32, device=dev), torch.empty(TRACE_CAP, dtype=torch.float32, device=dev)))
    out = torch.empty(cap, rows, D, dtype=torch.float32, device=dev)       # returned to the caller: always fresh
    t_out = torch.empty(cap, dtype=torch.float32, device=dev)
    t_host = np.ascontiguousarray(ts, dtype=np.float32)
    kern.mlp_solve(fm._desc, _rk_plan(solver), False, sign, atol, rtol, t_host, return_all_eval, y0, y1, k, out, t_out, cap,
                   fm.solve_workspace(kern), status, trace, rows, MAX_STEPS, comm)
    st = status.cpu().tolist()
    n_out, n_attempt, nfe, err = st[0], st[1], st[3], st[4]
    if err != 0:
        # 2: more accepted steps than the output buffer holds; 1: more than MAX_STEPS attempts (a legitimately long
        # solve, or a NaN error estimate).  Either way the host loop redoes the solve from x: it has the reference's
        # limits and raises on a NaN step size.
        return None
    _bump(inner.counters, nfe)
    _kernels.credit_work("mlp_solve", fm.flops_per_row * rows * nfe)
    _emit_trace(trace, n_attempt)
    _last = "persistent fused MLP integrator (tdyn_mlp_solve)" + (", error norm exchanged by NVLink peer stores" if comm else "")
    return t_out[:n_out], out[:n_out]


ADAPTIVE_TC = True        # adaptive solves of the 32-256-256-32 class: one cooperative launch (False: host-driven loop)
# The one-launch integrator builds every stage input in the prologue of its tile (x and up to 6 stage derivatives per row:
# exposed L2 / DRAM latency, ~20 % of the worker warps' time in the ncu capture), the host-driven loop runs the stage
# combination as its own HBM-bound launch.  Measured on B200 (benchmarks/k2_adaptive_ab.py, dopri5, 14 NFE): 4 096 rows 0.97
# vs 1.30 ms, 16 384 rows 1.02 vs 1.29 ms, 32 768 rows 1.28 vs 1.34 ms, 65 536 rows 1.87 vs 1.71 ms, 2^20 rows 17.7 vs 12.6 ms
# -- so the one-launch form takes the launch- / sync-bound sizes and the host loop the bandwidth-bound ones.  None: no limit
# (tests, A/B runs).
ADAPTIVE_TC_MAX_ROWS = 32768


def _try_adaptive_tc(inner, sign, x, ts, solver, atol, rtol, return_all_eval):
    """Adaptive solve of the 32-256-256-32 class in ONE launch of the tensor-core stage kernel
    (``tdyn_mlp_tc_solve_adaptive``: k1, init_step, every attempted step with its error norm and the controller)."""
    global _last
    from . import distributed as _dist
    fm, kern = inner.fm, inner.kern
    rows, D = x.shape
    n_t = len(ts)
    if not ADAPTIVE_TC or _dist.is_enabled() or n_t > MAX_T_SPAN or not (ts[-1] > ts[0]) or x.data_ptr() % 16 \
            or rows < 1 or (ADAPTIVE_TC_MAX_ROWS is not None and rows > ADAPTIVE_TC_MAX_ROWS):
        return None
    cap = n_t + (512 if return_all_eval else 0)
    dev = x.device
    x0 = x.clone()
    x1, k, status, trace, ws = fm.scratch(("adaptive_tc", rows, dev.index), lambda: (
        torch.empty_like(x), torch.empty(7 * rows * D, dtype=torch.float32, device=dev),
        torch.zeros(8, dtype=torch.int32, device=dev), torch.empty(TRACE_CAP, dtype=torch.float32, device=dev),
        kern.mlp_tc_adaptive_workspace()))
    out = torch.empty(cap, rows, D, dtype=torch.float32, device=dev)       # returned to the caller: always fresh
    t_out = torch.empty(cap, dtype=torch.float32, device=dev)
    t_host = np.ascontiguousarray(ts, dtype=np.float32)
    kern.mlp_tc_solve_adaptive(fm._desc, fm._img, _rk_plan(solver), sign, atol, rtol, t_host, return_all_eval, 0.0, x0, x1, k,
                               out, t_out, cap, ws, status, trace, rows, MAX_STEPS)
    st = status.cpu().tolist()
    n_out, n_attempt, nfe, err = st[0], st[1], st[3], st[4]
    if err != 0 or n_out == 0:
        return None                          # output buffer full / more than MAX_STEPS attempts / NaN: the host loop redoes it
    _bump(inner.counters, nfe)
    _kernels.credit_work("mlp_tc_adaptive", fm.flops_per_row * rows * nfe)
    _emit_trace(trace, n_attempt)
    _last = "persistent fused tcgen05 adaptive integrator (tdyn_mlp_tc_solve_adaptive)"
    return t_out[:n_out], out[:n_out]


def try_adjoint_interval(dyn, A, t_from, t_to, solver, atol, rtol):
    """One reverse-time interval of the backsolve adjoint (reference sensitivity.py:81-84) in ONE kernel launch.
    ``A`` = flat [x, lam, mu, lam_t]; integrates from ``t_from`` down to ``t_to``.  Returns the new flat A or None."""
    from . import distributed as _dist
    fa = getattr(dyn, "fused", None)
    if DISABLED or not PERSISTENT or not isinstance(fa, FusedAdjoint) or not _solver_ok(solver) or not A.is_contiguous():
        return None
    if getattr(dyn, "integral_loss", None) is not None:
        return None       # the -dg/dx term of an integral loss is a torch call: host-driven loop around the fused adjoint
    fm, kern = fa.fm, fa.kern
    if not kern.mlp_solve_supported(fm._desc, True):
        return None
    comm = None
    if _dist.is_enabled():
        if fm.n_params > 16384:
            return None
        comm = _dist.peer_comm(A.device)
        if comm is None:
            return None
    rows, D = dyn.x_shape
    dev = A.device
    y0 = A.clone()                           # the result is returned in y0 or y1: both fresh per call
    y1 = torch.empty_like(A)
    k, status, trace = fm.scratch(("adj", rows, dev.index), lambda: (
        torch.empty(7 * 2 * rows * D, dtype=torch.float32, device=dev), torch.zeros(8, dtype=torch.int32, device=dev),
        torch.empty(TRACE_CAP, dtype=torch.float32, device=dev)))
    t_host = np.array([-float(t_from), -float(t_to)], dtype=np.float32)     # negated span (odeint.py:54-58)
    kern.mlp_solve(fm._desc, _rk_plan(solver), True, -1.0, atol, rtol, t_host, False, y0, y1, k, None, None, 0,
                   fm.solve_workspace(kern), status, trace, rows, MAX_STEPS, comm)
    st = status.cpu().tolist()
    if st[4] != 0:
        return None                          # more than MAX_STEPS attempts: the host loop redoes the interval from A
    _bump(fa.counters, st[3])
    _kernels.credit_work("mlp_solve", 3.0 * fm.flops_per_row * rows * st[3])
    _emit_trace(trace, st[1])
    return y0 if st[5] == 0 else y1


def try_interp_interval(dyn, A, t_from, t_to, solver, atol, rtol):
    """One reverse-time interval of the INTERPOLATED adjoint (reference sensitivity.py:151-152) in ONE kernel launch.
    ``A`` = flat [lam, mu]; x(t) comes from ``dyn.spline`` (built by tdyn_spline_build).  Returns the new A or None."""
    from . import distributed as _dist
    fa = getattr(dyn, "fused", None)
    sp = getattr(dyn, "spline", None)
    if DISABLED or not PERSISTENT or not isinstance(fa, FusedAdjoint) or _dist.is_enabled() or not _solver_ok(solver) or not A.is_contiguous():
        return None
    if not all(hasattr(sp, a) for a in ("sol", "kd", "two_c", "three_d", "t")) or getattr(dyn, "integral_loss", None) is not None:
        return None
    fm, kern = fa.fm, fa.kern
    if not kern.mlp_solve_supported(fm._desc, True):
        return None
    rows, D = dyn.lam_shape
    dev = A.device
    if getattr(sp, "t_dev", None) is None:
        sp.t_dev = torch.from_numpy(np.ascontiguousarray(sp.t, dtype=np.float32)).to(dev)
    y0, y1 = A.clone(), torch.empty_like(A)
    k = torch.empty(7 * rows * D, dtype=torch.float32, device=dev)
    status = torch.zeros(8, dtype=torch.int32, device=dev)
    trace = torch.empty(TRACE_CAP, dtype=torch.float32, device=dev)
    kern.mlp_solve_interp(fm._desc, _rk_plan(solver), atol, rtol, -float(t_from), -float(t_to), y0, y1, k, sp.sol, sp.kd,
                          sp.two_c, sp.three_d, sp.t_dev, fm.solve_workspace(kern), status, trace, rows, MAX_STEPS)
    st = status.cpu().tolist()
    if st[4] != 0:
        return None                          # more than MAX_STEPS attempts: the host loop redoes the interval from A
    _bump(fa.counters, st[3])
    _kernels.credit_work("mlp_solve_interp", 3.0 * fm.flops_per_row * rows * st[3])
    _emit_trace(trace, st[1])
    return y0 if st[5] == 0 else y1


def _isclose_count(t, save_at: np.ndarray) -> int:
    return int(np.sum(np.abs(f32(t) - save_at) <= (1e-8 + 1e-5 * np.abs(save_at))))


def try_fixed(f, x, ts, sv, solver) -> Optional[List[torch.Tensor]]:
    """Fixed-step solve (euler / rk4) of a fusable MLP field entirely inside this library: one
    ``tdyn_mlp_tc_stage`` launch per RK stage (stage combination in its prologue, the solution update in the
    last stage's epilogue).  Returns the list of saved states or None when the generic path must be used."""
    global _last
    if DISABLED or not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2):
        return None
    if isinstance(f, FusedField):
        fm, counters = f.fm, f.counters
    else:
        got = _get_fused(f)
        if got is None:
            return None
        fm, counters = got
        fm.begin_solve()
    kern = _kernels.get(x)
    if not getattr(kern, "is_cuda", False) or x.shape[1] != fm.dims[0] or not fm.ready(kern):
        return None
    x = x if x.is_contiguous() else x.contiguous()
    p = solver.plan()
    n_stage = len(p.stages) + 1
    if SOLVE_KERNEL and x.data_ptr() % 16 == 0 and kern.mlp_tc_solve_supported(fm._desc, x.shape[0]):
        # the whole solve in ONE launch: dt and the saved time points follow the reference's loop (odeint.py:420-434)
        n_steps = len(ts) - 1
        dts, slot = np.empty(n_steps, dtype=np.float32), np.full(n_steps + 1, -1, dtype=np.int32)
        t, dt, n_saved = f32(ts[0]), f32(ts[1] - ts[0]), 0
        first = bool(_isclose_count(t, sv))
        for step in range(1, len(ts)):
            dts[step - 1] = dt
            t = f32(t + dt)
            if _isclose_count(t, sv):
                slot[step] = n_saved
                n_saved += 1
            if step < len(ts) - 1:
                dt = f32(ts[step + 1] - t)
        if slot[n_steps] >= 0:
            out = torch.empty(n_saved, *x.shape, dtype=torch.float32, device=x.device)       # returned to the caller: fresh
            xw, kbuf = fm.scratch(("solve_fixed", x.shape[0], n_stage, x.device.index), lambda: (
                torch.empty_like(x), torch.empty(n_stage, *x.shape, dtype=torch.float32, device=x.device)))
            kern.mlp_tc_solve_fixed(fm._desc, fm._img, _rk_plan(solver), x, xw, kbuf, out, dts, slot, n_stage, fm.flops_per_row)
            _bump(counters, n_stage * n_steps)
            _last = "persistent fused tcgen05 split-precision (TF32 + bf16 corrections) fixed-step integrator (tdyn_mlp_tc_solve_fixed)"
            return ([x] if first else []) + list(out.unbind(0))
    ks = [torch.empty_like(x) for _ in range(n_stage)]
    t, dt = f32(ts[0]), f32(ts[1] - ts[0])
    sol = [x] if _isclose_count(t, sv) else []
    spare, keep_x = None, True      # never recycle the caller's tensor or a saved state
    with torch.no_grad():
        for step in range(1, len(ts)):
            x_new = spare if spare is not None else torch.empty_like(x)
            for i in range(n_stage):
                last = i == n_stage - 1
                coefs = p.stages[i - 1].coefs if i > 0 else ()
                mode = p.stages[i - 1].mode if i > 0 else MODE_INNER
                kern.mlp_tc_stage(fm._desc, fm._img, ks[i], x, ks[:len(coefs)], coefs, mode, dt,
                                  x_new if last else None, p.bsol if last else (), fm.flops_per_row)
            t = f32(t + dt)
            saved = bool(_isclose_count(t, sv))
            if saved:
                sol.append(x_new)
            spare = None if keep_x else x
            x, keep_x = x_new, saved
            if step < len(ts) - 1:
                dt = f32(ts[step + 1] - t)
    _bump(counters, n_stage * (len(ts) - 1))
    _last = "fused tcgen05 split-precision (TF32 + bf16 corrections) MLP stage kernel (tdyn_mlp_tc_stage)"
    return sol
