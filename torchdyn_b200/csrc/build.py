"""Build libtdyn_b200.so in-tree with nvcc for sm_100a (no torch headers: the ABI is plain C).

    python torchdyn_b200/csrc/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtdyn_b200.so")
SOURCES = ["tdyn_generic.cu", "tdyn_mlp_tc.cu", "tdyn_mlp_ffma.cu", "tdyn_mlp_solve.cu", "tdyn_cnf_ffma.cu", "tdyn_wide_tc.cu", "tdyn_graph_loop.cu", "tdyn_peer.cu"]
HEADERS = ["tdyn_common.cuh", "tdyn_mlp_ffma.cuh", "tdyn_tc_ptx.cuh", os.path.join("..", "..", "include", "tdyn_b200.h")]
NVCC_FLAGS = (["-DTDYN_FFMA_FEAT_TILE=" + os.environ["TDYN_FFMA_FEAT_TILE"]] if os.environ.get("TDYN_FFMA_FEAT_TILE") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--fmad=true",  # parity-critical code uses explicit __f*_rn intrinsics; the MLP dot products may fuse
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    objs, procs = [], []
    for src in sources():                   # the translation units are independent: compile them side by side
        obj = src[:-3] + ".o"
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)     # (serial: ptxas -v output stays readable)
        else:
            procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lcuda", "-lpthread", "-ldl", "-lrt"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
