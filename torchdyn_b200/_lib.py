"""ctypes binding of libtdyn_b200.so (the sm_100a kernel library, C ABI in include/tdyn_b200.h).

There is no CPU fallback: if the shared library is missing or a kernel call fails, a RuntimeError is
raised.  The library is built in-tree by ``torchdyn_b200/csrc/build.py`` (also run by
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtdyn_b200.so")
ABI_VERSION = 15
MAX_STAGES = 7
MAX_LAYERS = 4
MODE_CHAIN, MODE_INNER = 0, 1
ACT_TANH, ACT_SOFTPLUS, ACT_RELU = 0, 1, 2

_lock = threading.Lock()
_lib = None

c_float_p = C.POINTER(C.c_float)
c_void_pp = C.POINTER(C.c_void_p)


class MlpDesc(C.Structure):
    """mirror of tdyn_mlp_t"""
    _fields_ = [("n_layers", C.c_int), ("dims", C.c_int * (MAX_LAYERS + 1)), ("act", C.c_int),
                ("w", C.c_void_p * MAX_LAYERS), ("b", C.c_void_p * MAX_LAYERS)]


class RkPlan(C.Structure):
    """mirror of tdyn_rk_plan_t"""
    _fields_ = [("n_extra", C.c_int), ("mode", C.c_int * 6), ("nk", C.c_int * 6), ("a", (C.c_float * MAX_STAGES) * 6),
                ("c", C.c_float * 6), ("n_sol", C.c_int), ("bsol", C.c_float * MAX_STAGES), ("n_err", C.c_int),
                ("berr", C.c_float * MAX_STAGES), ("order", C.c_int), ("safety", C.c_float), ("min_factor", C.c_float),
                ("max_factor", C.c_float)]


class Comm(C.Structure):
    """mirror of tdyn_comm_t"""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("epoch", C.c_uint), ("peer", C.c_void_p * 8)]


# name -> (restype, argtypes); every symbol include/tdyn_b200.h declares
SIGNATURES = {
    "tdyn_version": (C.c_int, []),
    "tdyn_probe_ffma": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "tdyn_last_error": (C.c_char_p, []),
    "tdyn_reduce_workspace_bytes": (C.c_int64, []),
    "tdyn_rk_combine": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp, c_float_p, C.c_int, C.c_int, C.c_float,
                                  C.c_int64, C.c_void_p]),
    "tdyn_rk_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_void_pp, c_float_p, C.c_int, c_float_p, C.c_int,
                                 C.c_float, C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "tdyn_init_norms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int64, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "tdyn_dense_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_void_pp, c_float_p, C.c_float, c_float_p,
                                  C.c_int64, C.c_void_p]),
    "tdyn_spline_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                    C.c_float, C.c_void_p]),
    "tdyn_spline_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64,
                                   C.c_void_p]),
    "tdyn_mlp_ffma_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "tdyn_mlp_ffma_workspace_bytes": (C.c_int64, [C.POINTER(MlpDesc)]),
    "tdyn_mlp_ffma_forward": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "tdyn_mlp_ffma_adjoint": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "tdyn_cnf_ffma_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "tdyn_cnf_ffma_forward": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "tdyn_cnf_ffma_adjoint": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "tdyn_mlp_solve_supported": (C.c_int, [C.POINTER(MlpDesc), C.c_int]),
    "tdyn_mlp_solve_workspace_bytes": (C.c_int64, [C.POINTER(MlpDesc)]),
    "tdyn_mlp_solve": (C.c_int, [C.POINTER(MlpDesc), C.POINTER(RkPlan), C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.POINTER(Comm), C.c_void_p]),
    "tdyn_comm_buffer_bytes": (C.c_int64, []),
    "tdyn_mlp_solve_interp": (C.c_int, [C.POINTER(MlpDesc), C.POINTER(RkPlan), C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int,
                                        C.c_void_p]),
    "tdyn_mlp_tc_prepared_bytes": (C.c_int64, [C.POINTER(MlpDesc)]),
    "tdyn_mlp_tc_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "tdyn_mlp_tc_prepare": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p]),
    "tdyn_mlp_tc_stage": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_void_pp,
                                    c_float_p, C.c_int, C.c_int, C.c_float, C.c_void_p, c_float_p, C.c_int, C.c_int64,
                                    C.c_void_p]),
    "tdyn_mlp_tc_forward_save": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_void_p]),
    "tdyn_mlp_tc_dgrad": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tdyn_mlp_tc_solve_supported": (C.c_int, [C.POINTER(MlpDesc), C.c_int64]),
    "tdyn_mlp_tc_solve_workspace_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "tdyn_mlp_tc_solve_fixed": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.POINTER(RkPlan), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "tdyn_mlp_tc_adaptive_workspace_bytes": (C.c_int64, []),
    "tdyn_mlp_tc_solve_adaptive": (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.POINTER(RkPlan), C.c_float, C.c_float, C.c_float,
                                             C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_int64, C.c_int, C.c_void_p]),
    "tdyn_lin_tc_supported": (C.c_int, [C.c_int, C.c_int]),
    "tdyn_lin_tc_image_bytes": (C.c_int64, [C.c_int, C.c_int]),
    "tdyn_lin_tc_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tdyn_lin_tc_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_float, C.c_int64, C.c_void_p]),
    "tdyn_wgrad_tc_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int64]),
    "tdyn_loop_ctrl_sizes": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "tdyn_rk_combine_dev": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp, c_float_p, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                      C.c_void_p]),
    "tdyn_rk_finish_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_void_pp, c_float_p, C.c_int, c_float_p, C.c_int,
                                     C.c_void_p, C.c_float, C.c_float, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tdyn_loop_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tdyn_loop_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tdyn_loop_commit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "tdyn_graph_while_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tdyn_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tdyn_graph_destroy": (C.c_int, [C.c_void_p]),
    "tdyn_peer_allreduce_buffer_bytes": (C.c_int64, [C.c_int64]),
    "tdyn_peer_allreduce": (C.c_int, [C.POINTER(Comm), C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]),
    "tdyn_wgrad_tc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p,
                                C.c_int64, C.c_int64, C.c_void_p]),
}


def load():
    """Load (once) and return the ctypes handle.  Raises RuntimeError if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"torchdyn_b200: CUDA kernel library not found at {LIB_PATH}. Build it with "
                "`python torchdyn_b200/csrc/build.py` (needs nvcc, targets sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype, fn.argtypes = res, args
        v = lib.tdyn_version()
        if v != ABI_VERSION:
            raise RuntimeError(f"torchdyn_b200: libtdyn_b200.so has ABI {v}, python expects {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().tdyn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"torchdyn_b200: {what} failed (code {rc}): {msg}")
