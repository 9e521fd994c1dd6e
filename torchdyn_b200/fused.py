"""Recognition of fusable MLP vector fields and dispatch to the fused kernels."""
from __future__ import annotations

DISABLED = False          # bench.py --force-generic / A-B parity runs: keep f a torch call
_last = None


def last_used():
    """Name of the fused path the most recent solve took (None = generic kernels, f via torch)."""
    return _last


def try_fixed(f, x, ts, sv, solver):
    global _last
    _last = None
    return None
