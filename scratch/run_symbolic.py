"""Time the symbolic phase (device + wall) at a given size. python scratch/run_symbolic.py [cells] [geometry cartesian|hex|perturbed]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import graft_import
g = graft_import.load(); L = g.libgraft
from helpers import build_problem
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 96
geom = sys.argv[2] if len(sys.argv) > 2 else "cartesian"
pr = build_problem((1, 1, 1), (cells,) * 3, 2, "boundary", lambda x: x[0] + x[1] + x[2], "sub")
kw = dict(geometry="cartesian") if geom == "cartesian" else dict(geometry="hex", perturb=g.vertex_perturbation(0.1, seed=1) if geom == "perturbed" else None)
assem = g.SparseMatrixAssembler(pr.U, pr.V, g.SubAssembledRows(), **kw)
form = g.Poisson(g.Measure(pr.trian, 4), source=1.0)
assem._set_form(form)
lib, comm = assem.comm.lib, assem.comm.handle
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter()
    L.check(lib.graft_symbolic(comm, 0, 0)); L.check(lib.graft_sync(comm))
    wall = (time.perf_counter() - t0) * 1e3
    print({"rep": rep, "cells": cells, "geometry": geom, "symbolic_ms_device": float(assem.timers()[0][L.T_SYMBOLIC]), "wall_ms": wall}, flush=True)
st = assem.stats()[0]
print(st)
assem.close()
