"""Thin Python face of the CUDA kernels: turns torch tensors into raw pointers + the current stream.

``get(x)`` returns the per-device ``CudaKernels`` object or raises -- there is deliberately no CPU
implementation in the product (tests inject a checker-backed double through ``set_test_backend``
to exercise the host logic without a GPU).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib

_launches = 0          # kernels launched by this library in this process (bench.py reports it)
_tls = threading.local()
_test_backend = None


def launches() -> int:
    return _launches


class KernelTimer:
    """Optional CUDA-event timing of every launch of this library (bench.py's roofline numbers): events are
    recorded on the launching stream right around the launch; ``summary()`` synchronises and aggregates."""

    def __init__(self):
        self.records = []   # (name, start_event, end_event, algorithmic_bytes_or_flops)
        self.credits = {}   # name -> algorithmic work only known after the launch (persistent solves: NFE from the status)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, work in self.records:
            a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["work"] += work
        for name, work in self.credits.items():
            if name in agg:
                agg[name]["work"] += work
        return agg


_timer: Optional[KernelTimer] = None


def set_timer(timer: Optional[KernelTimer]):
    global _timer
    _timer = timer


def credit_work(name: str, work: float):
    """Book algorithmic FLOPs of a launch after the fact (the persistent integrators report their NFE in a status word)."""
    if _timer is not None:
        _timer.credits[name] = _timer.credits.get(name, 0.0) + float(work)


class _Timed:
    __slots__ = ("name", "work", "e0")

    def __init__(self, name, work):
        self.name, self.work = name, work

    def __enter__(self):
        if _timer is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _timer is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _timer.records.append((self.name, self.e0, e1, self.work))
        return False


def set_test_backend(factory):
    """TESTS ONLY: ``factory(x) -> backend`` replaces the CUDA backend (None restores it)."""
    global _test_backend
    _test_backend = factory


def _fp(vals: Sequence[float]):
    return (C.c_float * len(vals))(*vals)


def _pp(tensors: Sequence[Tensor]):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda idx: torch.cuda.current_stream(idx).cuda_stream)


def _on_own_device(fn):
    """The C entry points launch on the calling thread's CURRENT device; a state on cuda:1 while cuda:0 is current is
    ordinary PyTorch usage, so every launching method switches to the instance's device for the duration of the call."""
    import functools

    @functools.wraps(fn)
    def guarded(self, *a, **k):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return guarded


def _guard_launchers(cls):
    skip = {"empty_like", "time_tensor"}
    for name, attr in list(vars(cls).items()):
        if callable(attr) and not name.startswith("_") and not isinstance(attr, staticmethod) and name not in skip:
            setattr(cls, name, _on_own_device(attr))
    return cls


@_guard_launchers
class CudaKernels:
    """One instance per (device, thread).  Owns the reduction scratch (device), its fp64 result and a
    pinned host mirror."""

    is_cuda = True

    def __init__(self, device: torch.device):
        self.lib = _lib.load()
        self.device = device
        nbytes = int(self.lib.tdyn_reduce_workspace_bytes())
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.red = torch.zeros(4, dtype=torch.float64, device=device)
        self.red_host = torch.zeros(4, dtype=torch.float64).pin_memory()

    # -- helpers ------------------------------------------------------------------------------
    def _stream(self):
        # (the raw handle of PyTorch's current stream on this device: ~0.3 us, against ~8 us for the Stream object)
        return _raw_stream(self.device.index)

    @staticmethod
    def _chk(t: Tensor, name: str):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise RuntimeError(f"torchdyn_b200: `{name}` must be a contiguous CUDA float32 tensor "
                               f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()})")

    def empty_like(self, x: Tensor) -> Tensor:
        return torch.empty_like(x, memory_format=torch.contiguous_format)

    def time_tensor(self, value, like: Tensor) -> Tensor:
        """shape-[1] time tensor on the state's device (what the reference hands to f)."""
        return torch.full((1,), float(value), dtype=like.dtype, device=like.device)

    # -- kernels ------------------------------------------------------------------------------
    def combine(self, y: Tensor, x: Optional[Tensor], ks: List[Tensor], coefs: Sequence[float], mode: int, dt: float):
        global _launches
        n = y.numel()
        if n == 0:
            return y
        for t in ks:
            self._chk(t, "k")
        if x is not None:
            self._chk(x, "x")
        with _Timed("rk_combine", 4.0 * n * (len(ks) + 1 + (x is not None))):
            rc = self.lib.tdyn_rk_combine(y.data_ptr(), _ptr(x), _pp(ks), _fp(coefs), len(ks), mode, float(dt), n,
                                          self._stream())
        _lib.check(rc, "tdyn_rk_combine")
        _launches += 1
        return y

    def finish(self, x_new: Tensor, err_out: Optional[Tensor], x: Tensor, ks: List[Tensor], bsol: Sequence[float],
               berr: Sequence[float], dt: float, atol: float, rtol: float, norm_n: int):
        """Launch K6; the fp64 sum of squared scaled errors lands in ``self.red[0]``."""
        global _launches
        self._chk(x, "x")
        for t in ks:
            self._chk(t, "k")
        with _Timed("rk_finish", 4.0 * x.numel() * (max(len(bsol), len(berr)) + 2 + (err_out is not None))):
            rc = self.lib.tdyn_rk_finish(x_new.data_ptr(), _ptr(err_out), x.data_ptr(), _pp(ks), _fp(bsol), len(bsol),
                                         _fp(berr), len(berr), float(dt), float(atol), float(rtol), x.numel(), int(norm_n),
                                         self.red.data_ptr(), self.ws.data_ptr(), self._stream())
        _lib.check(rc, "tdyn_rk_finish")
        _launches += 1

    # -- device-side loop (CUDA-graph WHILE, numerics/graph_loop.py): dt comes from the controller block on the device
    def combine_dev(self, y: Tensor, x: Optional[Tensor], ks: List[Tensor], coefs: Sequence[float], mode: int, dt_dev: int):
        global _launches
        rc = self.lib.tdyn_rk_combine_dev(y.data_ptr(), _ptr(x), _pp(ks), _fp(coefs), len(ks), mode, dt_dev, y.numel(),
                                          self._stream())
        _lib.check(rc, "tdyn_rk_combine_dev")
        _launches += 1
        return y

    def finish_dev(self, x_new: Tensor, x: Tensor, ks: List[Tensor], bsol, berr, dt_dev: int, atol: float, rtol: float,
                   norm_n: int):
        global _launches
        rc = self.lib.tdyn_rk_finish_dev(x_new.data_ptr(), None, x.data_ptr(), _pp(ks), _fp(bsol), len(bsol), _fp(berr),
                                         len(berr), dt_dev, float(atol), float(rtol), x.numel(), int(norm_n),
                                         self.red.data_ptr(), self.ws.data_ptr(), self._stream())
        _lib.check(rc, "tdyn_rk_finish_dev")
        _launches += 1

    def loop_begin(self, cf: Tensor, ci: Tensor, t_eval: Tensor):
        global _launches
        _lib.check(self.lib.tdyn_loop_begin(cf.data_ptr(), ci.data_ptr(), t_eval.data_ptr(), self._stream()), "tdyn_loop_begin")
        _launches += 1

    def loop_end(self, cf: Tensor, ci: Tensor, norm_count: int, t_eval: Tensor, t_out: Tensor, trace: Optional[Tensor]):
        global _launches
        _lib.check(self.lib.tdyn_loop_end(cf.data_ptr(), ci.data_ptr(), self.red.data_ptr(), int(norm_count), t_eval.data_ptr(),
                                          t_out.data_ptr(), _ptr(trace), self._stream()), "tdyn_loop_end")
        _launches += 1

    def loop_commit(self, ci: Tensor, x: Tensor, x_new: Tensor, k1: Tensor, k_last: Tensor, out: Tensor):
        global _launches
        _lib.check(self.lib.tdyn_loop_commit(ci.data_ptr(), x.data_ptr(), x_new.data_ptr(), k1.data_ptr(), k_last.data_ptr(),
                                             out.data_ptr(), x.numel(), self._stream()), "tdyn_loop_commit")
        _launches += 1

    def graph_while_create(self, raw_graph: int, keep_going_ptr: int) -> int:
        h = C.c_void_p()
        _lib.check(self.lib.tdyn_graph_while_create(raw_graph, keep_going_ptr, C.byref(h)), "tdyn_graph_while_create")
        return h.value

    def graph_launch(self, exec_handle: int):
        _lib.check(self.lib.tdyn_graph_launch(exec_handle, self._stream()), "tdyn_graph_launch")

    def graph_destroy(self, exec_handle: int):
        self.lib.tdyn_graph_destroy(exec_handle)

    def init_norms(self, x0: Tensor, f0: Tensor, f1: Optional[Tensor], atol: float, rtol: float):
        """Launch K9; sums land in ``self.red[0:3]``."""
        global _launches
        self._chk(x0, "x0"); self._chk(f0, "f0")
        if f1 is not None:
            self._chk(f1, "f1")
        rc = self.lib.tdyn_init_norms(x0.data_ptr(), f0.data_ptr(), _ptr(f1), float(atol), float(rtol), x0.numel(),
                                      self.red.data_ptr(), self.ws.data_ptr(), self._stream())
        _lib.check(rc, "tdyn_init_norms")
        _launches += 1

    def dense_eval(self, out: Tensor, x0: Tensor, x1: Tensor, ks: List[Tensor], bmid: Sequence[float], dt: float,
                   th: Sequence[float]):
        global _launches
        for t in (x0, x1, *ks):
            self._chk(t, "dense input")
        rc = self.lib.tdyn_dense_eval(out.data_ptr(), x0.data_ptr(), x1.data_ptr(), _pp(ks), _fp(bmid), float(dt),
                                      _fp(th), out.numel(), self._stream())
        _lib.check(rc, "tdyn_dense_eval")
        _launches += 1
        return out

    def spline_build(self, sol: Tensor, aux: Optional[Tensor], kd: Tensor, two_c: Tensor, three_d: Tensor, dt01: float):
        global _launches
        self._chk(sol, "sol")
        m = sol.shape[0]
        rc = self.lib.tdyn_spline_build(sol.data_ptr(), _ptr(aux), kd.data_ptr(), two_c.data_ptr(), three_d.data_ptr(),
                                        m, sol.numel() // m, float(dt01), self._stream())
        _lib.check(rc, "tdyn_spline_build")
        _launches += 1

    def spline_eval(self, out: Tensor, a: Tensor, b: Tensor, two_c: Tensor, three_d: Tensor, s: float):
        global _launches
        rc = self.lib.tdyn_spline_eval(out.data_ptr(), a.data_ptr(), b.data_ptr(), two_c.data_ptr(), three_d.data_ptr(),
                                       float(s), out.numel(), self._stream())
        _lib.check(rc, "tdyn_spline_eval")
        _launches += 1
        return out

    # -- fused MLP (tensor-core) path -----------------------------------------------------------
    def mlp_tc_supported(self, desc) -> bool:
        return bool(self.lib.tdyn_mlp_tc_supported(C.byref(desc)))

    def mlp_tc_prepare(self, desc) -> Tensor:
        global _launches
        nbytes = int(self.lib.tdyn_mlp_tc_prepared_bytes(C.byref(desc)))
        img = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.tdyn_mlp_tc_prepare(C.byref(desc), img.data_ptr(), self._stream()), "tdyn_mlp_tc_prepare")
        _launches += 1
        return img

    def mlp_tc_stage(self, desc, img: Tensor, k_out: Tensor, x: Tensor, ks: List[Tensor], coefs: Sequence[float],
                     mode: int, dt: float, x_new: Optional[Tensor], bsol: Sequence[float], flops_per_row: float):
        global _launches
        rows = x.shape[0]
        kp = _pp(ks) if ks else None
        cf = _fp(coefs) if len(coefs) else None
        bs = _fp(bsol) if len(bsol) else None
        with _Timed("mlp_tc_stage", flops_per_row * rows):
            rc = self.lib.tdyn_mlp_tc_stage(C.byref(desc), img.data_ptr(), k_out.data_ptr(), None, x.data_ptr(), kp, cf,
                                            len(coefs), mode, float(dt), _ptr(x_new), bs, len(bsol), rows, self._stream())
        _lib.check(rc, "tdyn_mlp_tc_stage")
        _launches += 1

    def mlp_tc_forward_save(self, desc, img: Tensor, out: Tensor, x: Tensor, scale: float, a1: Tensor, a2: Tensor,
                            flops_per_row: float):
        """out = scale * f(x) with both hidden activations written out (forward half of the adjoint dynamics)."""
        global _launches
        rows = x.numel() // 32
        with _Timed("mlp_tc_stage", flops_per_row * rows):
            rc = self.lib.tdyn_mlp_tc_forward_save(C.byref(desc), img.data_ptr(), out.data_ptr(), x.data_ptr(), float(scale),
                                                   a1.data_ptr(), a2.data_ptr(), rows, self._stream())
        _lib.check(rc, "tdyn_mlp_tc_forward_save")
        _launches += 1

    def mlp_tc_dgrad(self, desc_t, img_t: Tensor, out: Tensor, g: Tensor, scale: float, a2: Tensor, a1: Tensor, d2: Tensor,
                     d1: Tensor, flops_per_row: float):
        """Input-gradient chain of the 32-256-256-32 class in one launch (tdyn_mlp_tc_dgrad); d2 / d1 feed the wgrads."""
        global _launches
        rows = g.numel() // 32
        with _Timed("mlp_tc_stage", flops_per_row * rows):
            rc = self.lib.tdyn_mlp_tc_dgrad(C.byref(desc_t), img_t.data_ptr(), out.data_ptr(), g.data_ptr(), float(scale),
                                            a2.data_ptr(), a1.data_ptr(), d2.data_ptr(), d1.data_ptr(), rows, self._stream())
        _lib.check(rc, "tdyn_mlp_tc_dgrad")
        _launches += 1

    def mlp_tc_solve_supported(self, desc, rows: int) -> bool:
        return bool(self.lib.tdyn_mlp_tc_solve_supported(C.byref(desc), int(rows)))

    def mlp_tc_solve_fixed(self, desc, img: Tensor, plan, x: Tensor, xw: Tensor, k: Tensor, out: Tensor, dts, save_slot,
                           n_stages: int, flops_per_row: float):
        """Whole fixed-step solve in one launch (tdyn_mlp_tc_solve_fixed); dts / save_slot are host numpy arrays."""
        global _launches
        rows, n_steps = x.shape[0], int(dts.size)
        ws = torch.empty(int(self.lib.tdyn_mlp_tc_solve_workspace_bytes(n_steps, n_stages)), dtype=torch.uint8, device=self.device)
        with _Timed("mlp_tc_solve", flops_per_row * rows * n_steps * n_stages):
            rc = self.lib.tdyn_mlp_tc_solve_fixed(C.byref(desc), img.data_ptr(), C.byref(plan), x.data_ptr(), xw.data_ptr(),
                                                  k.data_ptr(), out.data_ptr(), dts.ctypes.data, save_slot.ctypes.data, n_steps,
                                                  rows, ws.data_ptr(), self._stream())
        _lib.check(rc, "tdyn_mlp_tc_solve_fixed")
        _launches += 1

    def probe_ffma_tflops(self, iters: int = 4096, reps: int = 5) -> float:
        """Sustained FP32 FMA rate of this GPU in TFLOP/s (register-operand FFMA chains on a full grid, CUDA events, best
        of `reps`): the measured denominator of the FFMA kernels' roofline."""
        out = torch.zeros(1, dtype=torch.float32, device=self.device)
        flops = C.c_double(0.0)
        best = 0.0
        for rep in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(self.lib.tdyn_probe_ffma(out.data_ptr(), int(iters), C.byref(flops), self._stream()), "tdyn_probe_ffma")
            e1.record()
            e1.synchronize()
            if rep > 0:
                best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best

    def mlp_tc_adaptive_workspace(self) -> Tensor:
        return torch.empty(int(self.lib.tdyn_mlp_tc_adaptive_workspace_bytes()), dtype=torch.uint8, device=self.device)

    def mlp_tc_solve_adaptive(self, desc, img: Tensor, plan, sign: float, atol: float, rtol: float, t_span, return_all: bool,
                              dt0: float, x0: Tensor, x1: Tensor, k: Tensor, out: Tensor, t_out: Tensor, out_cap: int,
                              ws: Tensor, status: Tensor, trace: Optional[Tensor], rows: int, max_steps: int):
        """Whole adaptive solve of the 32-256-256-32 class in one cooperative launch of the tensor-core stage kernel
        (tdyn_mlp_tc_solve_adaptive); results are read from `status` by the caller, who also books the work."""
        global _launches
        with _Timed("mlp_tc_adaptive", 0.0):
            rc = self.lib.tdyn_mlp_tc_solve_adaptive(C.byref(desc), img.data_ptr(), C.byref(plan), float(sign), float(atol),
                                                     float(rtol), t_span.ctypes.data, t_span.size, int(return_all), float(dt0),
                                                     x0.data_ptr(), x1.data_ptr(), k.data_ptr(), out.data_ptr(), t_out.data_ptr(),
                                                     int(out_cap), ws.data_ptr(), status.data_ptr(), _ptr(trace),
                                                     0 if trace is None else trace.numel(), int(rows), int(max_steps),
                                                     self._stream())
        _lib.check(rc, "tdyn_mlp_tc_solve_adaptive")
        _launches += 1

    # ---- layer-wise tensor-core kernels for wide MLPs (tdyn_lin_tc_*, tdyn_wgrad_tc)
    def lin_tc_supported(self, n_out: int, k_in: int) -> bool:
        return bool(self.lib.tdyn_lin_tc_supported(int(n_out), int(k_in)))

    def lin_tc_prepare(self, w: Tensor, transpose: bool) -> Tensor:
        """Weight image of one layer: B = W (forward) or B = W^T (input gradient)."""
        global _launches
        rows, cols = w.shape
        n, k = (cols, rows) if transpose else (rows, cols)
        img = torch.empty(int(self.lib.tdyn_lin_tc_image_bytes(n, k)), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.tdyn_lin_tc_prepare(w.data_ptr(), rows, cols, int(transpose), img.data_ptr(), self._stream()),
                   "tdyn_lin_tc_prepare")
        _launches += 1
        return img

    def lin_tc_apply(self, img: Tensor, n_out: int, k_in: int, inp: Tensor, bias: Optional[Tensor], saved: Optional[Tensor],
                     out: Tensor, act: int, mode: int, scale: float, rows: int):
        """out[r, n] = scale * epi(in[r, :] . B[n, :] + bias[n]); mode 0 linear, 1 activation, 2 times act'(saved)."""
        global _launches
        with _Timed("lin_tc", 2.0 * n_out * k_in * rows):
            rc = self.lib.tdyn_lin_tc_apply(img.data_ptr(), int(n_out), int(k_in), inp.data_ptr(), _ptr(bias), _ptr(saved),
                                            out.data_ptr(), int(act), int(mode), float(scale), int(rows), self._stream())
        _lib.check(rc, "tdyn_lin_tc_apply")
        _launches += 1

    def wgrad_tc_workspace(self, n_out: int, k_in: int, rows: int) -> Tensor:
        return torch.empty(int(self.lib.tdyn_wgrad_tc_workspace_bytes(int(n_out), int(k_in), int(rows))), dtype=torch.uint8,
                           device=self.device)

    def wgrad_tc(self, g: Tensor, n_out: int, inp: Tensor, k_in: int, dw: Tensor, db: Tensor, scale: float, ws: Tensor, rows: int):
        """dw[n, k] = scale * sum_r g[r, n] in[r, k]; db[n] = scale * sum_r g[r, n] (two launches: partials + reduce)."""
        global _launches
        with _Timed("wgrad_tc", 2.0 * n_out * k_in * rows):
            rc = self.lib.tdyn_wgrad_tc(g.data_ptr(), int(n_out), inp.data_ptr(), int(k_in), dw.data_ptr(), db.data_ptr(),
                                        float(scale), ws.data_ptr(), ws.numel(), int(rows), self._stream())
        _lib.check(rc, "tdyn_wgrad_tc")
        _launches += 2

    def mlp_ffma_supported(self, desc) -> bool:
        return bool(self.lib.tdyn_mlp_ffma_supported(C.byref(desc)))

    def mlp_ffma_workspace(self, desc) -> Tensor:
        return torch.zeros(int(self.lib.tdyn_mlp_ffma_workspace_bytes(C.byref(desc))), dtype=torch.uint8, device=self.device)

    def mlp_ffma_forward(self, desc, y: Tensor, out: Tensor, sign: float, flops_per_row: float):
        global _launches
        with _Timed("mlp_ffma_forward", flops_per_row * y.shape[0]):
            rc = self.lib.tdyn_mlp_ffma_forward(C.byref(desc), y.data_ptr(), out.data_ptr(), y.shape[0], float(sign),
                                                self._stream())
        _lib.check(rc, "tdyn_mlp_ffma_forward")
        _launches += 1
        return out

    def mlp_ffma_adjoint(self, desc, x: Tensor, lam: Tensor, dx: Tensor, dlam: Tensor, dmu: Tensor, ws: Tensor,
                         rows: int, sign: float, flops_per_row: float):
        """x, lam, dx, dlam: [rows*D] fp32 (possibly views into the flat adjoint vector); dmu: [P]."""
        global _launches
        with _Timed("mlp_ffma_adjoint", 3.0 * flops_per_row * rows):
            rc = self.lib.tdyn_mlp_ffma_adjoint(C.byref(desc), x.data_ptr(), lam.data_ptr(), dx.data_ptr(), dlam.data_ptr(),
                                                dmu.data_ptr(), ws.data_ptr(), int(rows), float(sign), self._stream())
        _lib.check(rc, "tdyn_mlp_ffma_adjoint")
        _launches += 1

    def mlp_solve_supported(self, desc, adjoint: bool) -> bool:
        return bool(self.lib.tdyn_mlp_solve_supported(C.byref(desc), int(adjoint)))

    def mlp_solve_workspace(self, desc) -> Tensor:
        return torch.zeros(int(self.lib.tdyn_mlp_solve_workspace_bytes(C.byref(desc))), dtype=torch.uint8, device=self.device)

    def mlp_solve(self, desc, plan, adjoint: bool, sign: float, atol: float, rtol: float, t_span, return_all: bool,
                  y0: Tensor, y1: Tensor, k: Tensor, out: Optional[Tensor], t_out: Optional[Tensor], out_cap: int, ws: Tensor,
                  status: Tensor, trace: Optional[Tensor], rows: int, max_steps: int, comm=None):
        """Whole odeint call in one cooperative kernel (tdyn_mlp_solve); results are read from `status` by the caller."""
        global _launches
        with _Timed("mlp_solve", 0.0):
            rc = self.lib.tdyn_mlp_solve(C.byref(desc), C.byref(plan), int(adjoint), float(sign), float(atol), float(rtol),
                                         t_span.ctypes.data, t_span.size, int(return_all), y0.data_ptr(), y1.data_ptr(),
                                         k.data_ptr(), _ptr(out), _ptr(t_out), int(out_cap), ws.data_ptr(), status.data_ptr(),
                                         _ptr(trace), 0 if trace is None else trace.numel(), int(rows), int(max_steps),
                                         None if comm is None else C.byref(comm), self._stream())
        _lib.check(rc, "tdyn_mlp_solve")
        _launches += 1

    def mlp_solve_interp(self, desc, plan, atol, rtol, t_start, t_end, y0, y1, k, sol, kd, two_c, three_d, knots, ws, status,
                         trace, rows, max_steps):
        """One interval of the interpolated adjoint in one cooperative kernel (tdyn_mlp_solve_interp)."""
        global _launches
        with _Timed("mlp_solve_interp", 0.0):
            rc = self.lib.tdyn_mlp_solve_interp(C.byref(desc), C.byref(plan), float(atol), float(rtol), float(t_start),
                                                float(t_end), y0.data_ptr(), y1.data_ptr(), k.data_ptr(), sol.data_ptr(),
                                                kd.data_ptr(), two_c.data_ptr(), three_d.data_ptr(), knots.data_ptr(),
                                                knots.numel(), ws.data_ptr(), status.data_ptr(), _ptr(trace),
                                                0 if trace is None else trace.numel(), int(rows), int(max_steps),
                                                self._stream())
        _lib.check(rc, "tdyn_mlp_solve_interp")
        _launches += 1

    def cnf_ffma_supported(self, desc) -> bool:
        return bool(self.lib.tdyn_cnf_ffma_supported(C.byref(desc)))

    def cnf_ffma_forward(self, desc, s: Tensor, eps: Tensor, out: Tensor, sign: float, flops_per_row: float):
        global _launches
        with _Timed("cnf_ffma_forward", 2.0 * flops_per_row * s.shape[0]):
            rc = self.lib.tdyn_cnf_ffma_forward(C.byref(desc), s.data_ptr(), eps.data_ptr(), out.data_ptr(), s.shape[0],
                                                float(sign), self._stream())
        _lib.check(rc, "tdyn_cnf_ffma_forward")
        _launches += 1
        return out

    def cnf_ffma_adjoint(self, desc, s: Tensor, lam: Tensor, eps: Tensor, ds: Tensor, dlam: Tensor, dmu: Tensor, ws: Tensor,
                         rows: int, sign: float, flops_per_row: float):
        global _launches
        with _Timed("cnf_ffma_adjoint", 6.0 * flops_per_row * rows):
            rc = self.lib.tdyn_cnf_ffma_adjoint(C.byref(desc), s.data_ptr(), lam.data_ptr(), eps.data_ptr(), ds.data_ptr(),
                                                dlam.data_ptr(), dmu.data_ptr(), ws.data_ptr(), int(rows), float(sign),
                                                self._stream())
        _lib.check(rc, "tdyn_cnf_ffma_adjoint")
        _launches += 1

    def read_reduction(self, n: int = 1):
        """Device -> pinned host read of the first ``n`` reduction results (the step's ONE host sync)."""
        self.red_host[:n].copy_(self.red[:n], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.red_host[:n].tolist()


def get(x: Tensor):
    """Kernel backend for the device of ``x``; raises when the CUDA path cannot run (no fallback)."""
    if _test_backend is not None:
        return _test_backend(x)
    if not isinstance(x, Tensor):
        raise NotImplementedError(f"{type(x)} is not supported as the state variable")
    if not x.is_cuda:
        raise RuntimeError("torchdyn_b200 integrates on B200 GPUs only: the state must be a CUDA tensor "
                           f"(got device {x.device}); there is no CPU fallback")
    if x.dtype != torch.float32:
        raise NotImplementedError(f"torchdyn_b200 kernels compute in float32; got state dtype {x.dtype}")
    cache = getattr(_tls, "cache", None)
    if cache is None:
        cache = _tls.cache = {}
    key = x.device.index if x.device.index is not None else torch.cuda.current_device()
    k = cache.get(key)
    if k is None:
        k = cache[key] = CudaKernels(torch.device("cuda", key))
    return k
