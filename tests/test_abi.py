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


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: the package (host mirror + CUDA sources), the header and the Julia shim must not
    import, link or mention it; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may."""
    pkg = os.path.join(ROOT, "gridapdistributed.jl_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        if os.sep + "build" in base or os.sep + "lib" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt):
                    offenders.append(os.path.join(base, f))
    for f in (os.path.join(ROOT, "include", "graft.h"), os.path.join(ROOT, "julia", "GraftAssembly.jl")):
        if re.search(r"\boracle\b", open(f).read()):
            offenders.append(f)
    assert not offenders, offenders
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py: the oracle only inside the CPU-baseline function
    uses = [m.start() for m in re.finditer(r"from oracle|import oracle", bench)]
    fn = bench.index("def cpu_reference_run")
    fn_end = bench.index("\ndef ", fn + 1)
    assert uses and all(fn < u < fn_end for u in uses)


def _abi_driver():
    import __graft_entry__ as entry

    return entry.build_abi_driver()


def test_abi_driver_links_against_the_library():
    """tests/abi_driver.c (the ccall sequence of julia/GraftAssembly.jl in plain C) compiles against include/graft.h alone and
    resolves every symbol it uses from libgraft.so; `--link` stops before the first GPU call."""
    import subprocess

    out = subprocess.run([_abi_driver(), "--link"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "link ok" in out.stdout, out.stdout + out.stderr


def test_julia_shim_binds_only_declared_symbols():
    """Every `:graft_*` symbol the Julia shim ccalls is declared in include/graft.h, and the shim covers hook 1 and the blocks."""
    txt = open(os.path.join(ROOT, "julia", "GraftAssembly.jl")).read()
    used = set(re.findall(r"\(:(graft_[a-z0-9_]+),\s*libgraft\)", txt))
    syms = set(header_symbols())
    assert used and used <= syms, used - syms
    for must in ("graft_scatter_cellmats", "graft_symbolic", "graft_numeric", "graft_spmv", "graft_csr_get", "graft_prange_get"):
        assert must in used
    # the C driver performs the same sequence
    drv = open(os.path.join(ROOT, "tests", "abi_driver.c")).read()
    drv_used = set(re.findall(r"\b(graft_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", drv, flags=re.S)))
    assert {"graft_scatter_cellmats", "graft_symbolic", "graft_numeric", "graft_spmv", "graft_cg", "graft_csr_get"} <= drv_used <= syms


@pytest.mark.gpu
def test_abi_driver_runs_the_shim_sequence_on_the_gpu():
    import subprocess

    for n in ("12", "3", "40"):
        out = subprocess.run([_abi_driver(), n], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "abi_driver: ok" in out.stdout, out.stdout + out.stderr
