"""CPU: the C-ABI library loads, exports every symbol ``include/rip_b200.h`` declares, and the ctypes mirrors of the
ABI structs have the compiled sizes.  No compute calls (no GPU here)."""

import ctypes as C
import os
import re

import pytest

from romanimpreprocess_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rip_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(rip_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_what_python_binds():
    assert declared_functions() == _lib.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/rip_b200.h but not exported by librip_b200.so"
    assert lib.rip_abi_version() == 1


def test_struct_sizes_match():
    lib = _lib.lib()
    for which, t in enumerate((_lib.RampSlice, _lib.RampPlan, _lib.CaldirDesc, _lib.L1L2Params, _lib.L2Out,
                               _lib.FwdParams)):  # fmt: skip
        assert lib.rip_struct_size(which) == C.sizeof(t), t.__name__
    assert lib.rip_struct_size(99) == -1


def test_errors_do_not_cross_the_boundary():
    """A failing call returns non-zero and leaves a message; here: no CUDA device / driver in the build container,
    or bad arguments on a GPU box."""
    lib = _lib.lib()
    rc = lib.rip_lin_eval(0, None, 7, None, 99, 10, 1, None, None)  # P out of range, bad dtype
    assert rc != 0
    assert b"rip_lin_eval" in lib.rip_last_error()
    with pytest.raises(_lib.RipError):
        _lib.check(rc)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librip_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()
