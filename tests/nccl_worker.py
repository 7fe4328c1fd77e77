"""Worker of tests/test_gpu_nccl.py: one part per rank / GPU over NCCL (the analogue of the reference's with_mpi mode).
Every rank assembles ITS part through libgraft.so with the NCCL communicator (ghost rows first, early exchange, merged
add) and checks it against the CPU oracle run on all parts in this process: index maps bit-exact, values to 1e-12;
then mul! (halo exchange over NCCL) and the CG solve against the oracle's gathered system."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    import graft_import
    import scipy.sparse.linalg as spla

    g = graft_import.load()
    from gpu_helpers import assert_same_prange
    from helpers import build_problem, gather_global, oracle_assemble

    u = lambda x: x[0] * x[1] + 2.0 * x[2]
    grid = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, (world, 1, 1))
    cases = [(grid, tuple(4 * p for p in grid[:2]) + (3 * grid[2],), "sub", "cartesian"),
             (grid[::-1], tuple(3 * p for p in grid[::-1]), "fully", "cartesian"),
             (grid, tuple(4 * p for p in grid), "sub", "perturbed"),        # general hexes: the fused sweep route under NCCL
             (grid, tuple(3 * p for p in grid), "fully", "perturbed")]
    pert = g.vertex_perturbation(0.1, seed=5)
    for parts, cells, strategy, geometry in cases:
        ref = build_problem(parts, cells, 2, "boundary", u, strategy)                       # all parts here (oracle side)
        operturb = (lambda m, lids, X: pert(m, m.vertex_coordinates())[m.cell_vertex_ids()[lids - 1] - 1]) if geometry == "perturbed" else None
        out, _ = oracle_assemble(ref, ("poisson",), source=1.0, perturb=operturb)
        pr = build_problem(parts, cells, 2, "boundary", u, strategy, backend=g.DistBackend())  # my part only
        st = g.FullyAssembledRows() if strategy == "fully" else g.SubAssembledRows()
        kw = dict(geometry="hex", perturb=pert) if geometry == "perturbed" else {}
        assem = g.SparseMatrixAssembler(pr.U, pr.V, st, device=local_rank, **kw)
        form = g.Poisson(g.Measure(pr.trian, 4), source=1.0)
        A, b = g.assemble_matrix_and_vector(form, assem)
        if geometry == "perturbed":
            assert assem.stats()[0]["path"] == "sumfact-gather"
        for rep in range(2):  # second pass: re-assembly (values only)
            p = out[rank]
            assert_same_prange(A.row_partition.indices[0], p["rows"], "rows")
            assert_same_prange(A.col_partition.indices[0], p["cols"], "cols")
            rowptr, colind, val = A.csr_arrays()[0]
            rp, ci, v = p["csr"]
            assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
            assert np.allclose(val, v, rtol=1e-12, atol=1e-12 * np.abs(v).max()), f"rank {rank}: values differ"
            assert np.allclose(b.vector_partition[0], p["b"], rtol=1e-12, atol=1e-12 * max(np.abs(p["b"]).max(initial=0.0), 1e-300))
            A, b = g.assemble_matrix_and_vector_b(A, b, form, assem)
        Ag, bg = gather_global(out)
        xg = np.random.default_rng(0).uniform(-1, 1, Ag.shape[0])
        x, y = g.pvector_on_cols(A, xg), g.pvector_on_rows(A)
        g.mul(y, A, x)
        ids = A.row_partition.indices[0]
        r = (Ag @ xg)[ids.l2g[: ids.own_length] - 1]
        assert np.allclose(y.vector_partition[0][: ids.own_length], r, rtol=1e-12, atol=1e-12 * np.abs(r).max())
        sol, (its, rr) = g.cg(A, b, rtol=1e-13, maxit=500)
        xref = spla.spsolve(Ag.tocsc(), bg)
        assert rr <= 1e-13
        assert np.allclose(sol.vector_partition[0][: ids.own_length], xref[ids.l2g[: ids.own_length] - 1], rtol=1e-10, atol=1e-10 * np.abs(xref).max())
        assem.close()
        if rank == 0:
            print(f"case parts={parts} cells={cells} {strategy} {geometry}: ok on {world} ranks", flush=True)
    # block system (Taylor-Hood Stokes): ghost rows of every block first, early exchange of the matrix values, right-hand side after
    # the step (numeric_fused_affine phases 1 / 2); against the oracle's block pipeline on all parts
    from helpers import stokes_oracle, stokes_problem

    for strategy in ("sub", "fully"):
        cells = tuple(2 * p for p in grid)
        ref = stokes_problem(grid, cells, strategy)
        out = stokes_oracle(ref, 0.7, 1.5)
        backend = g.DistBackend()
        pr = g.build_stokes_problem(grid, cells, strategy, backend=backend)
        st = g.FullyAssembledRows() if strategy == "fully" else g.SubAssembledRows()
        assem = g.SparseMatrixAssembler([pr.U, pr.P], [pr.V, pr.Q], st, device=local_rank)
        form = g.StokesTH(g.Measure(pr.trian, 4), nu=0.7, source=1.5)
        A, b = g.assemble_matrix_and_vector(form, assem)
        for rep in range(2):
            for i, j in ((0, 0), (0, 1), (1, 0)):
                p = out[i][j][rank]
                assert_same_prange(A[i][j].row_partition.indices[0], p["rows"], "rows")
                assert_same_prange(A[i][j].col_partition.indices[0], p["cols"], "cols")
                rowptr, colind, val = A[i][j].csr_arrays()[0]
                rp, ci, v = p["csr"]
                assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
                assert np.allclose(val, v, rtol=1e-12, atol=1e-12 * max(np.abs(v).max(initial=0.0), 1e-300)), f"rank {rank}: block ({i},{j}) differs"
                if j == 0:
                    assert np.allclose(b[i].vector_partition[0], p["b"], rtol=1e-12, atol=1e-12 * max(np.abs(p["b"]).max(initial=0.0), 1e-300))
            A, b = g.assemble_matrix_and_vector_b(A, b, form, assem)
        assem.close()
        if rank == 0:
            print(f"case Stokes parts={grid} cells={cells} {strategy}: ok on {world} ranks", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
