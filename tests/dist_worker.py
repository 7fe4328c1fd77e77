"""Worker of tests/test_dist_gloo.py: one part per rank over torch.distributed (gloo, CPU).  Builds the same
distributed problem (a) with the one-part-per-rank backend and (b) with all parts in this process (the debug
backend), and checks that this rank's part is identical: cell partition, global dof numbering (generate_gids,
reference FESpaces.jl:139-261: the three ghost-cell exchanges + the scan), Dirichlet values."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import graft_import

    g = graft_import.load()
    from helpers import build_problem

    u = lambda x: x[0] + 2 * x[1]
    for parts, cells, order in [((world, 1), (3 * world, 4), 2), ((1, world), (4, 2 * world + 1), 1)]:
        ref = build_problem(parts, cells, order, "boundary", u, "sub")                       # all parts here
        pr = build_problem(parts, cells, order, "boundary", u, "sub", backend=g.DistBackend())  # my part only
        assert len(pr.model.models) == 1 and pr.backend.parts_here == [rank + 1]
        a, b = pr.U.gids.indices[0], ref.U.gids.indices[rank]
        assert a.n_global == b.n_global and np.array_equal(a.l2g, b.l2g) and np.array_equal(a.l2o, b.l2o)
        ca, cb = pr.model.cell_gids.indices[0], ref.model.cell_gids.indices[rank]
        assert np.array_equal(ca.l2g, cb.l2g) and np.array_equal(ca.l2o, cb.l2o)
        assert np.array_equal(pr.U.spaces[0].cell_dof_ids, ref.U.spaces[rank].cell_dof_ids)
        assert np.allclose(pr.U.dirichlet_values[0], ref.U.dirichlet_values[rank])
        assert np.array_equal(pr.trian.cell_lids[0], ref.trian.cell_lids[rank])
    # round-2 input producers: periodic models (reference Geometry.jl:413-459: wrapped ghost cells in the partitioned periodic
    # direction), block spaces, and the boundary triangulation of the part
    ref = build_problem((world, 1), (3 * world, 4), 2, "boundary", u, "sub", isperiodic=(True, False))
    pr = build_problem((world, 1), (3 * world, 4), 2, "boundary", u, "sub", isperiodic=(True, False), backend=g.DistBackend())
    a, b = pr.U.gids.indices[0], ref.U.gids.indices[rank]
    assert a.n_global == b.n_global and np.array_equal(a.l2g, b.l2g) and np.array_equal(a.l2o, b.l2o)
    assert np.array_equal(pr.U.spaces[0].cell_dof_ids, ref.U.spaces[rank].cell_dof_ids)
    for tags in ("boundary", [6]):
        Ga, Gb = g.Boundary(pr.model, tags=tags), g.Boundary(ref.model, tags=tags)
        assert np.array_equal(Ga.cell_lids[0], Gb.cell_lids[rank]) and np.array_equal(Ga.lfaces[0], Gb.lfaces[rank])
    sref = g.build_stokes_problem((world, 1), (2 * world, 2), "sub")
    spr = g.build_stokes_problem((world, 1), (2 * world, 2), "sub", backend=g.DistBackend())
    for fa, fb in ((spr.U, sref.U), (spr.P, sref.P)):
        a, b = fa.gids.indices[0], fb.gids.indices[rank]
        assert a.n_global == b.n_global and np.array_equal(a.l2g, b.l2g) and np.array_equal(a.l2o, b.l2o)
        assert np.array_equal(fa.spaces[0].cell_dof_ids, fb.spaces[rank].cell_dof_ids)
    # backend primitives
    be = g.DistBackend()
    assert be.scan_exclusive([rank + 1]) == [sum(range(1, rank + 1))]
    assert be.reduction([rank + 1]) == [world * (world + 1) // 2]
    rcv = be.exchange([{q: np.arange(3, dtype=np.int64) + 10 * (rank + 1) for q in range(1, world + 1) if q != rank + 1}])
    for src, data in rcv[0].items():
        assert np.array_equal(data, np.arange(3) + 10 * src)
    be.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
