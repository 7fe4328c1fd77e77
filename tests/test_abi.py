"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/graft.h declares; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import graft_import

g = graft_import.load()
L = g.libgraft
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "graft.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(graft_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/graft.h but not exported by libgraft.so"
        assert s in L.SIGNATURES, f"{s} has no ctypes prototype"
    assert sorted(L.SIGNATURES) == syms


def test_version():
    assert L.load().graft_version() >= 100


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    comm = L.c_vp()
    st = L.load().graft_comm_create_local(1, C.byref(comm))
    assert st != 0
    assert b"no CPU fallback" in L.load().graft_last_error()
