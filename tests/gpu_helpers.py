"""Drivers shared by the GPU parity tests: run the product path (libgraft.so through the host mirror)
and the CPU oracle on the same problem, compare index maps bit-exactly and values to 1e-12."""
import numpy as np

from helpers import build_problem, g, oracle_assemble

FORMS = {"poisson": (g.Poisson, ("poisson",)), "mass": (g.Mass, ("mass",))}


def graft_assemble(pr, form="poisson", source=None, geometry="cartesian", perturb=None, index_base=0, extra_cellvec=None, params=()):
    strategy = g.FullyAssembledRows() if pr.strategy == "fully" else g.SubAssembledRows()
    assem = g.SparseMatrixAssembler(pr.U, pr.V, strategy, index_base=index_base, geometry=geometry, perturb=perturb)
    dΩ = g.Measure(pr.trian, 2 * pr.order)
    if form == "elasticity":
        f = g.LinearElasticity(dΩ, params[0], params[1], source=source, extra_cellvec=extra_cellvec)
    else:
        f = FORMS[form][0](dΩ, source=source, extra_cellvec=extra_cellvec)
    A, b = g.assemble_matrix_and_vector(f, assem)
    return assem, f, A, b


def assert_same_prange(ours, ref, what):
    assert ours.n_global == ref["n"], what
    nown = len(ref["own_to_local"])
    assert np.array_equal(ours.own_to_global, ref["l2g"][:nown]), f"{what}: own_to_global differs"
    assert np.array_equal(ours.ghost_to_global, ref["l2g"][nown:]), f"{what}: ghost_to_global differs"
    assert np.array_equal(ours.ghost_to_owner, ref["l2o"][nown:]), f"{what}: ghost_to_owner differs"


def assert_matches_oracle(A, b, out, tol=1e-12, check_b=True):
    """Sparsity pattern / local-to-global / ghost maps bit-exact; entries within tol relative,
    normwise per row (BASELINE.json north_star)."""
    csr = A.csr_arrays()
    base = A.assem.index_base
    for k, p in enumerate(out):
        assert_same_prange(A.row_partition.indices[k], p["rows"], f"part {k + 1} rows")
        assert_same_prange(A.col_partition.indices[k], p["cols"], f"part {k + 1} cols")
        rowptr, colind, val = csr[k]
        rp, ci, v = p["csr"]
        assert np.array_equal(rowptr - base, rp), f"part {k + 1}: rowptr differs"
        assert np.array_equal(colind - base, ci), f"part {k + 1}: colind differs"
        rid = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        err2 = np.zeros(len(rp) - 1); nrm2 = np.zeros(len(rp) - 1)
        np.add.at(err2, rid, (val - v) ** 2)
        np.add.at(nrm2, rid, v ** 2)
        scale = np.sqrt(nrm2.max()) if len(nrm2) else 1.0
        bad = np.sqrt(err2) > tol * np.maximum(np.sqrt(nrm2), 1e-300) + 1e-15 * scale
        assert not bad.any(), f"part {k + 1}: {bad.sum()} rows differ, max rel {np.sqrt(err2 / np.maximum(nrm2, 1e-300)).max():.3e}"
        if check_b and p["b"] is not None:
            bo = b.vector_partition[k]
            ref_rows = p["brows"] if p["brows"] is not None else p["rows"]
            assert_same_prange(b.index_partition.indices[k], ref_rows, f"part {k + 1} rows of b")
            assert np.allclose(bo, p["b"], rtol=tol, atol=tol * max(np.abs(p["b"]).max(initial=0.0), 1e-300)), f"part {k + 1}: b differs"


__all__ = ["build_problem", "oracle_assemble", "graft_assemble", "assert_matches_oracle", "assert_same_prange"]
