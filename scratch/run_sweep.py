"""Time the numeric step of one geometry route (device resident). python scratch/run_sweep.py [cells] [steps] [perturb 0/1]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import graft_import
g = graft_import.load(); L = g.libgraft
from helpers import build_problem
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pert = int(sys.argv[3]) if len(sys.argv) > 3 else 1
what = int(sys.argv[4]) if len(sys.argv) > 4 else 3
pr = build_problem((1, 1, 1), (cells,) * 3, 2, "boundary", lambda x: x[0] + x[1] + x[2], "sub")
rng = np.random.default_rng(0)
def perturb(m, xyz):
    ijk = m.vertex_multi_index(); n = np.asarray(m.ncells_local)
    mask = np.all((ijk > 0) & (ijk < n[None, :]), axis=1).astype(np.float64)
    return xyz + rng.uniform(-0.1, 0.1, xyz.shape) * np.asarray(m.h)[None, :] * mask[:, None]
assem = g.SparseMatrixAssembler(pr.U, pr.V, g.SubAssembledRows(), geometry="hex", perturb=perturb if pert else None)
form = g.Poisson(g.Measure(pr.trian, 4), source=1.0)
assem._set_form(form); assem._symbolic(form)
lib, comm = assem.comm.lib, assem.comm.handle
for _ in range(2):
    L.check(lib.graft_numeric(comm, what))
L.check(lib.graft_sync(comm))
ts = []
for _ in range(steps):
    L.check(lib.graft_numeric(comm, what)); L.check(lib.graft_sync(comm))
    ts.append(assem.timers()[0][L.T_NUMERIC]); ph = (float(assem.timers()[0][L.T_INTEGRATE]), float(assem.timers()[0][L.T_SCATTER]))
st = assem.stats()[0]
print({"what": what, "cells": cells, "route": st["path"], "ms": float(np.min(ts)), "ms_mean": float(np.mean(ts)), "nnz": st["nnz"], "all_ms": [round(float(t), 2) for t in ts], "integrate_scatter_ms_last": ph,
       "GBps_alg": (8 * st["nnz"] + 8 * st["nrows"] + 108 * st["ncells"] + 24 * (cells + 1) ** 3) / (np.min(ts) * 1e-3) / 1e9})
assem.close()
