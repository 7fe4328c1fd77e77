"""ctypes driver of oracle/assembly_ref.c (CPU ORACLE -- test infrastructure / CPU baseline, NOT product code).

``assemble_poisson_q2(pr)`` runs the C restatement of the reference's single-part path on every part of a
problem built by tests/helpers.build_problem -- one OS thread per part (the C call releases the GIL), like one
MPI rank per part in the reference's with_mpi mode -- and returns per part (rowptr, colind, vals, b)."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libassembly_ref.so")


class RefCSR(C.Structure):
    _fields_ = [("m", C.c_int64), ("nnz", C.c_int64), ("ncoo", C.c_int64), ("rowptr", C.POINTER(C.c_int64)),
                ("colind", C.POINTER(C.c_int64)), ("vals", C.POINTER(C.c_double)), ("b", C.POINTER(C.c_double))]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.ref_assemble_poisson_q2.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_double,
                                                 C.POINTER(RefCSR)]
        _lib.ref_assemble_poisson_q2.restype = C.c_int
        _lib.ref_csr_free.argtypes = [C.POINTER(RefCSR)]
    return _lib


def part_inputs(pr, k):
    """(cell vertex coordinates, tensor index of the local dofs, cell_dof_ids, nfree, Dirichlet values) of part k."""
    m, s = pr.model.models[k], pr.U.spaces[k]
    lids = pr.trian.cell_lids[k]
    X = np.ascontiguousarray(m.vertex_coordinates()[m.cell_vertex_ids()[lids - 1] - 1], dtype=np.float64)
    tix = np.ascontiguousarray(np.rint(np.asarray(s.ref_nodes) * pr.order), dtype=np.int32)
    ids = np.ascontiguousarray(s.cell_dof_ids[lids - 1], dtype=np.int32)
    dv = np.ascontiguousarray(pr.U.dirichlet_values[k], dtype=np.float64)
    if len(dv) == 0:
        dv = np.zeros(1)
    return X, tix, ids, int(s.num_free_dofs), dv


def run_part(inp, source=1.0, keep=True):
    lib = load()
    X, tix, ids, nfree, dv = inp
    out = RefCSR()
    st = lib.ref_assemble_poisson_q2(len(ids), X.ctypes.data, tix.ctypes.data, ids.ctypes.data, nfree, dv.ctypes.data, float(source),
                                     C.byref(out))
    if st != 0:
        raise MemoryError("oracle/assembly_ref.c: allocation failed")
    res = (out.m, out.nnz, out.ncoo)
    if keep:
        res = (np.ctypeslib.as_array(out.rowptr, (out.m + 1,)).copy(), np.ctypeslib.as_array(out.colind, (max(out.nnz, 1),))[: out.nnz].copy(),
               np.ctypeslib.as_array(out.vals, (max(out.nnz, 1),))[: out.nnz].copy(), np.ctypeslib.as_array(out.b, (max(out.m, 1),))[: out.m].copy())
    lib.ref_csr_free(C.byref(out))
    return res


def assemble_poisson_q2(pr, source=1.0, threads=None, keep=True, inputs=None):
    inputs = inputs if inputs is not None else [part_inputs(pr, k) for k in range(len(pr.model.models))]
    with ThreadPoolExecutor(max_workers=threads or len(inputs)) as ex:
        return list(ex.map(lambda i: run_part(i, source, keep), inputs))
