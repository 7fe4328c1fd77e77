import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_helpers import build_problem, graft_assemble, oracle_assemble
cells = tuple(int(c) for c in sys.argv[1].split(",")) if len(sys.argv) > 1 else (4, 3, 3)
src = float(sys.argv[2]) if len(sys.argv) > 2 else -1.5
u = (lambda x: x[0] * x[1] + 2.0 * x[2] + 0.25) if len(sys.argv) <= 3 else None
pr = build_problem((1, 1, 1), cells, 2, "boundary", u, "sub")
out, _ = oracle_assemble(pr, ("poisson",), source=src)
assem, f, A, b = graft_assemble(pr, "poisson", source=src, geometry="hex")
print(assem.stats()[0]["path"])
rp, ci, v = out[0]["csr"]
val = A.csr_arrays()[0][2]
print("A max err", np.abs(val - v).max(), "of", np.abs(v).max())
bo = b.vector_partition[0]; br = out[0]["b"]
err = np.abs(bo - br)
print("b max err", err.max(), "of", np.abs(br).max(), "nbad", (err > 1e-10).sum(), "of", len(br))
bad = np.flatnonzero(err > 1e-10)[:20]
print(bad, bo[bad], br[bad])
