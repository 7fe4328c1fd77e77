"""The C restatement of the reference path (oracle/assembly_ref.c, the CPU baseline of bench.py) against the
numpy oracle: identical local FE-space CSR pattern, values to 1e-13."""
import numpy as np
import pytest

from helpers import build_problem, oracle_dofs
from oracle import assembly_oracle as orc
from oracle import c_oracle


@pytest.mark.parametrize("cells", [(3, 3, 3), (4, 2, 3)])
def test_c_oracle_matches_numpy_oracle(cells):
    u = lambda x: x[0] - 2 * x[1] + 0.5 * x[2]
    pr = build_problem((1, 1, 1), cells, 2, "boundary", u, "sub")
    rowptr, colind, vals, b = c_oracle.assemble_poisson_q2(pr, source=1.5)[0]
    m, s = pr.model.models[0], pr.U.spaces[0]
    lids = pr.trian.cell_lids[0]
    X = m.vertex_coordinates()[m.cell_vertex_ids()[lids - 1] - 1]
    K, F = orc.integrate_cells(("poisson",), X, s.ref_nodes, 2, 1, 4, 1.5)
    ids = s.cell_dof_ids[lids - 1]
    F = orc.lift_dirichlet(K, F, ids, pr.U.dirichlet_values[0])
    I, J, V, bb, _ = orc.numeric_loop(ids, ids, K, F, s.num_free_dofs, None)
    rp, ci, v = orc.coo_to_csr(I, J, V, s.num_free_dofs, s.num_free_dofs)
    assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
    assert np.allclose(vals, v, rtol=1e-13, atol=1e-14 * np.abs(v).max())
    assert np.allclose(b, bb, rtol=1e-12, atol=1e-14)
    assert len(vals) == np.prod([8 * c - 9 for c in cells])  # nnz closed form (8N-9)^D, SURVEY 8c


def test_c_oracle_parts_in_threads():
    pr = build_problem((2, 1, 1), (4, 2, 2), 2, "boundary", None, "sub")
    res = c_oracle.assemble_poisson_q2(pr, source=1.0, keep=False)
    assert len(res) == 2 and all(r[1] > 0 for r in res)
