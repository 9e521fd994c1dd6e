"""Import the REAL reference (torchdyn 1.0.6) for oracle validation / golden generation.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only where the
reference tree exists (``$TORCHDYN_REF`` or ``/root/reference``) -- i.e. in the
build container, never on the GPU box.

The reference imports three third-party packages that are absent here
(SURVEY.md 8c / Appendix D):

* ``torchcde``  (numerics/utils.py:21, numerics/sensitivity.py:16)
* ``torchsde``  (core/problems.py:15, core/neuralde.py:24, numerics/sdeint.py:6)
* ``pytorch_lightning`` (core/neuralde.py:20)

Placeholders are registered in ``sys.modules`` before the import.  ``torchcde``'s
two names used on the hot path (``natural_cubic_coeffs``, ``CubicSpline``) are
bound to the oracle's restatement of the published torchcde 0.2.x algorithm
(oracle/erk_oracle.py, Appendix B of SURVEY.md) so the reference's interpolated
adjoint can execute; everything else on the path is the unmodified reference.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

_REF = None


def reference_root():
    for cand in (os.environ.get("TORCHDYN_REF"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "torchdyn")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def load():
    """Return the imported reference ``torchdyn`` module (cached)."""
    global _REF
    if _REF is not None:
        return _REF
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set $TORCHDYN_REF); it only exists in the build container")

    import torch.nn as nn
    from . import erk_oracle as eo

    if "torchcde" not in sys.modules:
        m = types.ModuleType("torchcde")

        class CubicSpline:  # same call surface as torchcde.CubicSpline for the calls in sensitivity.py:126-132
            def __init__(self, coeffs, t):
                self._sp = eo.NaturalCubicSpline.from_coeffs(coeffs, t)

            def evaluate(self, t):
                return self._sp.evaluate(t)

        def natural_cubic_coeffs(x, t=None):
            return eo.natural_cubic_coeffs(x, t)

        def _absent(*a, **k):
            raise NotImplementedError("torchcde is not installed; this helper is off the hot path")

        m.CubicSpline = CubicSpline
        m.natural_cubic_coeffs = natural_cubic_coeffs
        m.hermite_cubic_coefficients_with_backward_differences = _absent
        sys.modules["torchcde"] = m

    if "torchsde" not in sys.modules:
        m = types.ModuleType("torchsde")
        b = types.ModuleType("torchsde._brownian")

        class BaseBrownian:  # noqa: D401 - placeholder
            pass

        class BrownianInterval:
            pass

        b.BaseBrownian, b.BrownianInterval = BaseBrownian, BrownianInterval
        m._brownian = b
        m.BrownianInterval = BrownianInterval
        sys.modules["torchsde"] = m
        sys.modules["torchsde._brownian"] = b

    if "pytorch_lightning" not in sys.modules:
        m = types.ModuleType("pytorch_lightning")
        m.LightningModule = nn.Module
        sys.modules["pytorch_lightning"] = m

    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import torchdyn  # noqa: E402  (the reference)
        import torchdyn.numerics.odeint  # noqa: F401
        import torchdyn.numerics.sensitivity  # noqa: F401
        import torchdyn.core  # noqa: F401
    _REF = torchdyn
    return _REF


class StepTrace:
    """Record (t, dt, error_ratio) for every attempted adaptive step of the reference by wrapping
    its public hooks from the outside: the two adaptive steppers' ``step``
    (t, dt of the attempt), ``adapt_step`` (error_ratio) and ``_adaptive_odeint`` (call boundaries)."""

    def __init__(self):
        self.records = []  # one dict per adaptive odeint call
        self._cur = None

    def __enter__(self):
        load()
        import importlib
        ro = importlib.import_module("torchdyn.numerics.odeint")   # the package re-exports a function of the same name
        ro = sys.modules["torchdyn.numerics.odeint"]
        so = sys.modules["torchdyn.numerics.solvers.ode"]
        self._ro, self._so = ro, so
        self._orig_adapt = ro.adapt_step
        self._orig_adaptive = ro._adaptive_odeint
        self._orig_steps = {cls: cls.step for cls in (so.DormandPrince45, so.Tsitouras45)}
        trace = self

        def adapt_wrapped(dt, error_ratio, *a, **k):
            if trace._cur is not None:
                trace._cur["ratio"].append(float(error_ratio))
            return trace._orig_adapt(dt, error_ratio, *a, **k)

        def adaptive_wrapped(f, k1, x, dt, t_span, *a, **k):
            trace._cur = {"t0": float(t_span[0]), "T": float(t_span[-1]), "dt0": float(dt),
                          "t": [], "dt": [], "ratio": []}
            try:
                return trace._orig_adaptive(f, k1, x, dt, t_span, *a, **k)
            finally:
                trace.records.append(trace._cur)
                trace._cur = None

        def make_step(orig):
            def step(self_, f, x, t, dt, k1=None, args=None):
                if trace._cur is not None:
                    trace._cur["t"].append(float(t))
                    trace._cur["dt"].append(float(dt))
                return orig(self_, f, x, t, dt, k1=k1, args=args)
            return step

        ro.adapt_step = adapt_wrapped
        ro._adaptive_odeint = adaptive_wrapped
        for cls, orig in self._orig_steps.items():
            cls.step = make_step(orig)
        return self

    def __exit__(self, *exc):
        self._ro.adapt_step = self._orig_adapt
        self._ro._adaptive_odeint = self._orig_adaptive
        for cls, orig in self._orig_steps.items():
            cls.step = orig
        return False
