"""The C-ABI library loads and exports every symbol include/tdyn_b200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

from conftest import ROOT

from torchdyn_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tdyn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tdyn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with python torchdyn_b200/csrc/build.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tdyn_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.tdyn_version() == _lib.ABI_VERSION
    assert lib.tdyn_reduce_workspace_bytes() > 0
    # argument validation happens before any CUDA call: usable without a GPU
    rc = lib.tdyn_rk_combine(None, None, None, None, 1, 0, 0.1, 8, None)
    assert rc != 0 and b"null" in lib.tdyn_last_error()


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout
