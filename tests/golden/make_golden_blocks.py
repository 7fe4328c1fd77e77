"""Golden fixture of the block pipeline: Taylor-Hood Stokes on 4x4 cells, (2,2) parts, both
blocks rows, generated from the CPU oracle (oracle.create_from_nz_blocks).  Same purpose as make_golden.py.

    python tests/golden/make_golden_blocks.py        # rewrites tests/golden/stokes_2d_22_sub.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import graft_import  # noqa: E402
from helpers import Problem, cell_coords  # noqa: E402
from oracle import assembly_oracle as orc  # noqa: E402

g = graft_import.load()
NAME = "stokes_2d_22_sub"
NU, SRC = 0.7, 1.5
BLOCKS = ((0, 0), (0, 1), (1, 0))


def build(parts=(2, 2), cells=(4, 4), strategy="sub"):
    D = len(cells)
    pr = Problem()
    pr.backend = g.DebugBackend(int(np.prod(parts)))
    pr.model = g.CartesianDiscreteModel(pr.backend, parts, sum(([0.0, 1.0] for _ in cells), []), cells)
    pr.V = g.TestFESpace(pr.model, g.ReferenceFE("lagrangian", float, 2, ncomp=D), dirichlet_tags="boundary")
    pr.Q = g.TestFESpace(pr.model, g.ReferenceFE("lagrangian", float, 1), dirichlet_tags=None)
    pr.U = g.TrialFESpace(lambda x: np.stack([x[(d + 1) % D] * (1.0 - x[d]) + 0.25 * d for d in range(D)]), pr.V)
    pr.P = g.TrialFESpace(None, pr.Q)
    pr.strategy, pr.D = strategy, D
    pr.trian = g.Triangulation(g.FullyAssembledRows(), pr.model) if strategy == "fully" else g.Triangulation(pr.model)
    return pr


def oracle_blocks(pr, nu=NU, source=SRC):
    nf = 2
    spaces = [pr.U, pr.P]
    dofs = [[orc.local_indices(i.n_global, i.part, i.l2g, i.l2o) for i in sp.gids.indices] for sp in spaces]
    P = len(pr.model.models)
    I = [[[None] * P for _ in range(nf)] for _ in range(nf)]
    J = [[[None] * P for _ in range(nf)] for _ in range(nf)]
    V = [[[None] * P for _ in range(nf)] for _ in range(nf)]
    B = [[None] * P for _ in range(nf)]
    T = [[None] * P for _ in range(nf)]
    for k, m in enumerate(pr.model.models):
        lids = pr.trian.cell_lids[k]
        X = cell_coords(m, lids)
        su, sp_ = pr.U.spaces[k], pr.P.spaces[k]
        Kuu, Kup, Kpu, Fu = orc.integrate_stokes_cells(X, su.ref_nodes, sp_.ref_nodes, 2, 1, 4, nu, source)
        idu, idp = su.cell_dof_ids[lids - 1], sp_.cell_dof_ids[lids - 1]
        Fp = np.zeros((len(lids), idp.shape[1]))
        Fu = orc.lift_dirichlet(Kuu, Fu, idu, pr.U.dirichlet_values[k])
        Fu = orc.lift_dirichlet(Kup, Fu, idp, pr.P.dirichlet_values[k])
        Fp = orc.lift_dirichlet(Kpu, Fp, idu, pr.U.dirichlet_values[k])
        masks = [(d[k]["l2o"] != d[k]["part"]) if pr.strategy == "fully" else None for d in dofs]
        I[0][0][k], J[0][0][k], V[0][0][k], B[0][k], T[0][k] = orc.numeric_loop(idu, idu, Kuu, Fu, su.num_free_dofs, masks[0])
        I[0][1][k], J[0][1][k], V[0][1][k], _, _ = orc.numeric_loop(idu, idp, Kup, None, su.num_free_dofs, masks[0])
        I[1][0][k], J[1][0][k], V[1][0][k], B[1][k], T[1][k] = orc.numeric_loop(idp, idu, Kpu, Fp, sp_.num_free_dofs, masks[1])
    I[1][1] = J[1][1] = V[1][1] = None
    return orc.create_from_nz_blocks(pr.strategy, I, J, V, B, T, dofs, dofs)


def run():
    out = oracle_blocks(build())
    d = {}
    for (i, j) in BLOCKS:
        for k, p in enumerate(out[i][j]):
            pre = f"b{i}{j}_p{k}_"
            d[pre + "rows_l2g"] = p["rows"]["l2g"]; d[pre + "rows_l2o"] = p["rows"]["l2o"]
            d[pre + "rows_nown"] = np.int64(len(p["rows"]["own_to_local"]))
            d[pre + "cols_l2g"] = p["cols"]["l2g"]; d[pre + "cols_l2o"] = p["cols"]["l2o"]
            d[pre + "cols_nown"] = np.int64(len(p["cols"]["own_to_local"]))
            d[pre + "rowptr"], d[pre + "colind"], d[pre + "vals"] = p["csr"]
            if p["b"] is not None:
                d[pre + "b"] = p["b"]
    d["nparts"] = np.int64(len(out[0][0]))
    return d


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, NAME + ".npz"), **run())
    print("wrote", NAME)
