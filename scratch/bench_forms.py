"""Side measurement (not the contract bench): numeric re-assembly of the vector-valued / block configurations
(BASELINE.json configs 4 and 5) on one GPU.   python scratch/bench_forms.py elasticity 48 [cartesian|hex|perturb]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import graft_import

g = graft_import.load()
L = g.libgraft
from helpers import build_problem

form_name, n = sys.argv[1], int(sys.argv[2])
geom = sys.argv[3] if len(sys.argv) > 3 else "cartesian"
steps = 5
t0 = time.time()
kw = {}
if geom == "perturb":
    rng = np.random.default_rng(0)

    def perturb(m, xyz):
        ijk = m.vertex_multi_index()
        inside = np.all((ijk > 0) & (ijk < np.asarray(m.ncells_local)[None, :]), axis=1).astype(np.float64)
        return xyz + rng.uniform(-0.1, 0.1, xyz.shape) * np.asarray(m.h)[None, :] * inside[:, None]

    kw["perturb"] = perturb
geometry = "cartesian" if geom == "cartesian" else "hex"
backend = g.DebugBackend(1)
if form_name == "elasticity":
    pr = build_problem((1, 1, 1), (n, n, n), 2, "boundary", lambda x: np.stack([x[0], x[1], x[2]]), "sub", ncomp=3, backend=backend)
    assem = g.SparseMatrixAssembler(pr.U, pr.V, g.SubAssembledRows(), geometry=geometry, **kw)
    form = g.LinearElasticity(g.Measure(pr.trian, 4), 1.0, 1.0, source=1.0)
    blocks = [(0, 0)]
else:
    model = g.CartesianDiscreteModel(backend, (1, 1, 1), [0, 1, 0, 1, 0, 1], (n, n, n))
    V = g.TestFESpace(model, g.ReferenceFE("lagrangian", float, 2, ncomp=3), dirichlet_tags="boundary")
    Q = g.TestFESpace(model, g.ReferenceFE("lagrangian", float, 1), dirichlet_tags=None)
    U = g.TrialFESpace(lambda x: np.stack([x[1], x[2], x[0]]), V)
    P = g.TrialFESpace(None, Q)
    assem = g.SparseMatrixAssembler([U, P], [V, Q], g.SubAssembledRows(), geometry=geometry, **kw)
    form = g.StokesTH(g.Measure(g.Triangulation(model), 4), nu=1.0, source=1.0)
    blocks = [(0, 0), (0, 1), (1, 0)]
t_setup = time.time() - t0
lib, comm, ctx = assem.comm.lib, assem.comm.handle, assem.comm.ctxs[0]
assem._set_form(form)
t0 = time.time()
assem._symbolic(form)
t_sym = time.time() - t0
for _ in range(2):
    L.check(lib.graft_numeric(comm, 3))
L.check(lib.graft_sync(comm))
L.check(lib.graft_mark(comm, 0))
for _ in range(steps):
    L.check(lib.graft_numeric(comm, 3))
L.check(lib.graft_mark(comm, 1))
el = C.c_double()
L.check(lib.graft_elapsed(ctx, 0, 1, C.byref(el)))
ms = el.value / steps
nnz = 0
for bi, bj in blocks:
    m_, n_, z = L.c_i64(), L.c_i64(), L.c_i64()
    L.check(lib.graft_csr_query(ctx, bi, bj, C.byref(m_), C.byref(n_), C.byref(z)))
    nnz += z.value
st = assem.stats()[0]
tm = assem.timers()[0]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"form": form_name, "cells": n ** 3, "geometry": geom, "route": st["path"], "nnz": nnz, "ms_per_step": ms,
                  "nnz_per_s": nnz / (ms * 1e-3), "cells_per_s": n ** 3 / (ms * 1e-3), "values_GBps": 8 * nnz / (ms * 1e-3) / 1e9,
                  "frac_of_hbm_peak_values_only": 8 * nnz / (ms * 1e-3) / 1e9 / peak,
                  "phase_ms": {"integrate": float(tm[L.T_INTEGRATE]), "scatter": float(tm[L.T_SCATTER])},
                  "symbolic_s": t_sym, "host_setup_s": t_setup, "device_bytes": st["bytes_dev"]}))
assem.close()
