"""Pin the CPU oracle with every known-answer test the reference holds for the path."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import build_problem, g, gather_global, l2_error, neumann_cellvec_2d, oracle_assemble
from oracle import assembly_oracle as orc


def test_1d_q2_element_matrices():
    # SURVEY 8c analytic KAT: Gridap node order (left, right, mid)
    h = 0.7
    X = np.array([[[0.0], [h]]])
    ref = np.array([[0.0], [1.0], [0.5]])
    K, _ = orc.integrate_cells(("poisson",), X, ref, 2, 1, 4)
    M, _ = orc.integrate_cells(("mass",), X, ref, 2, 1, 4)
    assert np.allclose(K[0], np.array([[7, 1, -8], [1, 7, -8], [-8, -8, 16]]) / (3 * h), rtol=1e-13)
    assert np.allclose(M[0], np.array([[4, -1, 2], [-1, 4, 2], [2, 2, 16]]) * h / 30, rtol=1e-13)


def test_3d_cartesian_cell_matrix_is_kronecker_sum():
    from helpers import g
    hx, hy, hz = 0.5, 0.25, 2.0
    poly = g.NCube(3)
    ref = poly.q2_ref_nodes()
    X = (poly.vertex_coords * np.array([hx, hy, hz]))[None].astype(float)
    K, _ = orc.integrate_cells(("poisson",), X, ref, 2, 1, 4)
    k1 = lambda h: np.array([[7, -8, 1], [-8, 16, -8], [1, -8, 7]]) / (3 * h)  # tensor order (0, 1/2, 1)
    m1 = lambda h: np.array([[4, 2, -1], [2, 16, 2], [-1, 2, 4]]) * h / 30
    tix = np.rint(ref * 2).astype(int)
    Kx, Ky, Kz, Mx, My, Mz = k1(hx), k1(hy), k1(hz), m1(hx), m1(hy), m1(hz)
    a, b, c = tix[:, 0], tix[:, 1], tix[:, 2]
    ref_K = (Kx[np.ix_(a, a)] * My[np.ix_(b, b)] * Mz[np.ix_(c, c)] + Mx[np.ix_(a, a)] * Ky[np.ix_(b, b)] * Mz[np.ix_(c, c)]
             + Mx[np.ix_(a, a)] * My[np.ix_(b, b)] * Kz[np.ix_(c, c)])
    assert np.allclose(K[0], ref_K, rtol=1e-12, atol=1e-14)
    assert np.abs(K[0].sum(1)).max() < 1e-12  # row sums of the unconstrained Laplacian vanish


@pytest.mark.parametrize("strategy", ["sub", "fully"])
def test_poisson_config1_solution(strategy):
    # reference test/PoissonTests.jl:14-45: 4x4 cells on (0,4)^2, (2,2) parts, Q2, u=(x+y)^2
    u = lambda x: (x[0] + x[1]) ** 2
    pr = build_problem((2, 2), (4, 4), 2, [1, 2, 3, 5, 7], u, strategy, domain=[0, 4, 0, 4])
    assert sum(i.own_length for i in pr.U.gids.indices) == 64
    gN = lambda x, n: 2 * (x[0] + x[1]) * (n[0] + n[1])
    out, _ = oracle_assemble(pr, ("poisson",), source=-4.0, extra_cellvec=neumann_cellvec_2d(pr, gN))
    for p in out:  # test/FESpacesTests.jl:29-31
        assert len(p["csr"][0]) - 1 == len(p["rows"]["l2g"])
        assert p["csr"][1].max() < len(p["cols"]["l2g"])
    A, b = gather_global(out)
    x = spla.spsolve(A.tocsc(), b)
    assert l2_error(pr, x, u) < 1e-9


@pytest.mark.parametrize("strategy", ["sub", "fully"])
@pytest.mark.parametrize("parts", [(2, 2), (1, 1), (2, 1)])
def test_rhs_sum_equals_length(strategy, parts):
    # reference test/FESpacesTests.jl:61-70,174-181: l = ∫ 1*v, Q1, h=1, Dirichlet "boundary":
    # every free (interior) node integrates to exactly 1
    pr = build_problem(parts, (4, 4), 1, "boundary", None, strategy, domain=[0, 4, 0, 4])
    out, _ = oracle_assemble(pr, ("poisson",), source=1.0)
    _, b = gather_global(out)
    assert len(b) == 9 and abs(b.sum() - len(b)) < 1e-12


def test_fully_equals_sub_3d():
    u = lambda x: x[0] + 2 * x[1] - x[2]
    res = {}
    for st in ("sub", "fully"):
        pr = build_problem((2, 2, 2), (4, 4, 4), 2, "boundary", u, st)
        out, _ = oracle_assemble(pr, ("poisson",), source=1.0)
        res[st] = gather_global(out)
    dA = abs(res["sub"][0] - res["fully"][0]).max()
    assert dA < 1e-12 and np.abs(res["sub"][1] - res["fully"][1]).max() < 1e-12


def test_nnz_closed_forms_single_part():
    # nnz = (8N-9)^D, COO count = (9N-10)^D for Q2 with full Dirichlet boundary
    N = 4
    pr = build_problem((1, 1, 1), (N, N, N), 2, "boundary", None, "sub")
    out, KF = oracle_assemble(pr)
    assert len(out[0]["csr"][1]) == (8 * N - 9) ** 3
    ids = pr.U.spaces[0].cell_dof_ids
    assert int(((ids > 0).sum(1) ** 2).sum()) == (9 * N - 10) ** 3


def test_mul_alpha_beta():
    # reference test/BlockPartitionedArraysTests.jl:70-77: y <- alpha*(A x) + beta*y
    pr = build_problem((2, 2), (6, 6), 2, "boundary", None, "sub")
    out, _ = oracle_assemble(pr)
    A, _ = gather_global(out)
    rng = np.random.default_rng(0)
    xg = rng.uniform(-1, 1, A.shape[0])
    xs = [xg[p["cols"]["l2g"] - 1] for p in out]
    ys = [rng.uniform(-1, 1, len(p["rows"]["l2g"])) for p in out]
    yo = orc.mul(out, xs, 2.0, -0.5, ys)
    for p, y, y0 in zip(out, yo, ys):
        nown = len(p["rows"]["own_to_local"])
        ref = 2.0 * (A @ xg)[p["rows"]["l2g"][:nown] - 1] - 0.5 * y0[:nown]
        assert np.allclose(y, ref, rtol=1e-13, atol=1e-13)


def test_stokes_cell_blocks_kats():
    """Analytic properties of the Taylor-Hood cell blocks: Kpu = Kup^T, the Laplacian rows sum to zero, and a constant
    velocity is divergence free (Kpu applied to a constant vector vanishes)."""
    D = 3
    model = g.CartesianDiscreteModel(g.DebugBackend(1), (1, 1, 1), [0, 2, 0, 1, 0, 3], (2, 1, 2))
    V = g.TestFESpace(model, g.ReferenceFE("lagrangian", float, 2, ncomp=D), dirichlet_tags=None)
    Q = g.TestFESpace(model, g.ReferenceFE("lagrangian", float, 1), dirichlet_tags=None)
    m = model.models[0]
    X = m.vertex_coordinates()[m.cell_vertex_ids() - 1]
    Kuu, Kup, Kpu, Fu = orc.integrate_stokes_cells(X, V.spaces[0].ref_nodes, Q.spaces[0].ref_nodes, 2, 1, 4, 0.5, 2.0)
    assert np.allclose(Kpu, np.transpose(Kup, (0, 2, 1)), atol=1e-14)
    assert np.abs(Kuu.sum(axis=2)).max() < 1e-12
    assert np.abs(Kpu @ np.ones(Kpu.shape[2])).max() < 1e-13
    # source: sum of the cell vector of one component = f * |cell|
    nds = Kuu.shape[1] // D
    assert np.allclose(Fu[:, :nds].sum(axis=1), 2.0 * (1.0 * 1.0 * 1.5))


def test_block_pipeline_reduces_to_single_field():
    """create_from_nz_blocks with one field reproduces create_from_nz bit for bit (both strategies)."""
    for strategy in ("sub", "fully"):
        pr = build_problem((2, 2), (4, 4), 2, "boundary", lambda x: x[0] + x[1], strategy)
        out, _ = oracle_assemble(pr, ("poisson",), source=1.0)
        dofs = [orc.local_indices(i.n_global, i.part, i.l2g, i.l2o) for i in pr.U.gids.indices]
        I, J, Vv, B, T = [], [], [], [], []
        from helpers import cell_coords
        for k, (m, s) in enumerate(zip(pr.model.models, pr.U.spaces)):
            lids = pr.trian.cell_lids[k]
            K, F = orc.integrate_cells(("poisson",), cell_coords(m, lids), s.ref_nodes, 2, 1, 4, 1.0)
            ids = s.cell_dof_ids[lids - 1]
            F = orc.lift_dirichlet(K, F, ids, pr.U.dirichlet_values[k])
            mask = (dofs[k]["l2o"] != dofs[k]["part"]) if strategy == "fully" else None
            i, j, v, b, t = orc.numeric_loop(ids, ids, K, F, s.num_free_dofs, mask)
            I.append(i); J.append(j); Vv.append(v); B.append(b); T.append(t)
        blk = orc.create_from_nz_blocks(strategy, [[I]], [[J]], [[Vv]], [B], [T], [dofs], [dofs])[0][0]
        for a, c in zip(out, blk):
            for u, w in zip(a["csr"], c["csr"]):
                assert np.array_equal(u, w)
            assert np.array_equal(a["rows"]["l2g"], c["rows"]["l2g"]) and np.array_equal(a["cols"]["l2g"], c["cols"]["l2g"])
            assert np.array_equal(a["b"], c["b"])


def test_state_dependent_forms_jacobian_is_the_derivative_of_the_residual():
    """Oracle self-consistency for the forms of reference test/PLaplacianTests.jl:24-31: K(uh) = d r(uh) / d uh (central
    differences), on a general (non-affine) quad cell; and r(u) = 0 for the linear exact solution with f = 0."""
    from oracle import assembly_oracle as orc
    import graft_import
    g = graft_import.load()
    ref = g.NCube(2).q2_ref_nodes()
    X = np.array([[[0.0, 0.0], [1.1, 0.1], [-0.1, 0.9], [1.0, 1.2]]])
    rng = np.random.default_rng(0)
    uc = rng.uniform(-1, 1, (1, 9))
    K, F = orc.integrate_cells(("plaplacian", uc), X, ref, 2, 1, 4, None)
    eps = 1e-6
    for j in range(9):
        up, um = uc.copy(), uc.copy()
        up[0, j] += eps; um[0, j] -= eps
        Fp = orc.integrate_cells(("plaplacian", up), X, ref, 2, 1, 4, None)[1]
        Fm = orc.integrate_cells(("plaplacian", um), X, ref, 2, 1, 4, None)[1]
        assert np.allclose((Fp - Fm)[0] / (2 * eps), K[0, :, j], rtol=1e-6, atol=1e-8)
    # linear u: sigma(grad u) is constant, its divergence vanishes: interior residual entries are zero (Q2 cell-interior node = 8)
    Xa = np.array([[[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]])
    lin = (ref[:, 0] + ref[:, 1])[None, :]
    Fl = orc.integrate_cells(("plaplacian", lin), Xa, ref, 2, 1, 4, None)[1]
    interior = int(np.flatnonzero(np.all(np.isclose(ref, 0.5), axis=1))[0])
    assert abs(Fl[0, interior]) < 1e-13


@pytest.mark.parametrize("parts", [(1, 1), (1, 2), (2, 1), (2, 2)])
def test_dof_count_kat_nonperiodic(parts):
    """reference test/sequential/PeriodicBCsTests.jl:7-20, non-periodic column: Q1 on 4x4 cells without Dirichlet tags has 25
    free dofs on every part grid -- pins generate_gids (FESpaces.jl:139-261), the producer of every index map."""
    backend = g.DebugBackend(int(np.prod(parts)))
    model = g.CartesianDiscreteModel(backend, parts, [0, 1, 0, 1], (4, 4))
    V = g.FESpace(model, g.ReferenceFE("lagrangian", float, 1))
    ids = V.gids.indices
    assert all(i.n_global == 25 for i in ids)
    owned = np.concatenate([i.l2g[i.l2o == i.part] for i in ids])
    assert len(owned) == 25 and np.array_equal(np.sort(owned), np.arange(1, 26))   # every gid owned exactly once
    # owners: a dof belongs to the highest part among the cells around it (FESpaces.jl:154-156)
    for i in ids:
        assert np.all((i.l2o >= 1) & (i.l2o <= int(np.prod(parts))))


def test_measure_of_the_domain_kat():
    """reference test/CellDataTests.jl:48: sum(∫(1)dΩ) ≈ 16 on (0,4)^2 with 4x4 cells, (2,2) parts (owned cells only): through
    the oracle's quadrature, sum_cells sum_q w_q |det J_q| with degree 1."""
    pr = build_problem((2, 2), (4, 4), 1, None, None, "sub", domain=[0, 4, 0, 4])
    from helpers import cell_coords
    tot = 0.0
    for k, m in enumerate(pr.model.models):
        X = cell_coords(m, pr.trian.cell_lids[k])
        phi, dphi, w, xi = orc.reference_tables(2, 1, pr.U.spaces[k].ref_nodes, 1)
        N, dN = orc.geometry_tables(2, xi)
        J = np.einsum("cvd,qva->cqda", X, dN)
        tot += float(np.einsum("q,cq->", w, np.abs(np.linalg.det(J))))
    assert abs(tot - 16.0) < 1e-12


@pytest.mark.parametrize("strategy", ["sub", "fully"])
def test_stokes_solution_kat(strategy):
    """Solution-level KAT of the block pipeline, after reference test/MultiFieldTests.jl:27-57,69-76 (domain (0,4)^2, 4x4 cells,
    (2,2) parts, k = 2, u = ((x+y)^2, (x-y)^2), p = x+y - mean, f = -Δu + ∇p, g = ∇·u, BlockMultiFieldStyle): with the
    Taylor-Hood pair Q2^2/Q1 the exact solution lies in the FE space, so the discrete solution reproduces it (the reference
    checks l2 errors < 1e-9).  Pressure: continuous Q1 here (the reference uses P1-discontinuous + zero mean), fixed by its mean."""
    from helpers import cell_coords, stokes_oracle, stokes_problem
    import scipy.sparse as sp
    u = lambda x: np.stack([(x[0] + x[1]) ** 2, (x[0] - x[1]) ** 2])
    pfun = lambda x: x[0] + x[1] - 4.0
    pr = stokes_problem((2, 2), (4, 4), strategy, ufun=u, domain=[0, 4, 0, 4])
    # l((v,q)) = ∫ v⋅f - q*g with f = (-4 + 1, -4 + 1) constant and g = 4 y: the q*g term as pressure cell vectors
    extra_p = []
    for k, m in enumerate(pr.model.models):
        lids = pr.trian.cell_lids[k]
        X = cell_coords(m, lids)
        phip, _, w, xi = orc.reference_tables(2, 1, pr.P.spaces[k].ref_nodes, 4)
        N, dN = orc.geometry_tables(2, xi)
        xq = np.einsum("qv,cvd->cqd", N, X)
        wd = w[None, :] * np.abs(np.linalg.det(np.einsum("cvd,qva->cqda", X, dN)))
        extra_p.append(-np.einsum("cq,qi,cq->ci", wd, phip, 4.0 * xq[:, :, 1]))
    out = stokes_oracle(pr, 1.0, -3.0, extra_p=extra_p)
    nu_, np_ = pr.U.gids.indices[0].n_global, pr.P.gids.indices[0].n_global
    def gather_block(parts_out):
        nr, nc = parts_out[0]["rows"]["n"], parts_out[0]["cols"]["n"]
        rows, cols, vals, b = [], [], [], np.zeros(nr)
        for p in parts_out:
            rowptr, colind, val = p["csr"]
            nown = len(p["rows"]["own_to_local"])
            rid = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
            keep = rid < nown
            rows.append(p["rows"]["l2g"][rid[keep]] - 1); cols.append(p["cols"]["l2g"][colind[keep]] - 1); vals.append(val[keep])
            if p["b"] is not None:
                b[p["rows"]["l2g"][:nown] - 1] += p["b"][:nown]
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nr, nc)), b

    blocks = {}
    for (i, j) in ((0, 0), (0, 1), (1, 0)):
        blocks[(i, j)], b = gather_block(out[i][j])
        if j == 0:
            blocks[("b", i)] = b
    K = sp.bmat([[blocks[(0, 0)], blocks[(0, 1)]], [blocks[(1, 0)], None]], format="csr")
    rhs = np.concatenate([blocks[("b", 0)], blocks[("b", 1)]])
    # exact nodal values
    ue, pe = np.zeros(nu_), np.zeros(np_)
    for k in range(len(pr.model.models)):
        su, sq = pr.U.spaces[k], pr.P.spaces[k]
        iu, ip = pr.U.gids.indices[k], pr.P.gids.indices[k]
        vals = u(su.free_dof_coords.T)
        ue[iu.l2g - 1] = vals[su.free_dof_comp, np.arange(len(su.free_dof_comp))]
        pe[ip.l2g - 1] = pfun(sq.free_dof_coords.T)
    xe = np.concatenate([ue, pe])
    assert np.abs(K @ xe - rhs).max() < 1e-10 * np.abs(rhs).max()       # the interpolant solves the discrete system
    # and the system determines it (up to the pressure constant): solve with one pressure dof pinned
    keep = np.arange(nu_ + np_ - 1)
    x = np.zeros(nu_ + np_)
    x[-1] = xe[-1]
    x[keep] = spla.spsolve(K[keep][:, keep].tocsc(), rhs[keep] - K[keep][:, [nu_ + np_ - 1]].toarray().ravel() * xe[-1])
    assert np.abs(x - xe).max() < 1e-9


@pytest.mark.parametrize("parts", [(1, 1), (1, 2), (2, 1), (2, 2)])
@pytest.mark.parametrize("isperiodic,ndofs", [((False, False), 25), ((False, True), 20), ((True, False), 20), ((True, True), 16)])
def test_dof_count_kat_periodic(parts, isperiodic, ndofs):
    """reference test/sequential/PeriodicBCsTests.jl:7-20: ndofss = [25,20,20,16] for the four periodicity patterns on every
    part grid (periodic models: reference Geometry.jl:413-459)."""
    backend = g.DebugBackend(int(np.prod(parts)))
    model = g.CartesianDiscreteModel(backend, parts, [0, 1, 0, 1], (4, 4), isperiodic=isperiodic)
    V = g.FESpace(model, g.ReferenceFE("lagrangian", float, 1))
    assert all(i.n_global == ndofs for i in V.gids.indices)
    owned = np.concatenate([i.l2g[i.l2o == i.part] for i in V.gids.indices])
    assert np.array_equal(np.sort(owned), np.arange(1, ndofs + 1))


def _periodic_problem(parts, strategy="sub", cells=(20, 20)):
    from helpers import Problem
    u = lambda x: np.sin(x[1] + np.pi / 6) * x[0]
    pr = Problem()
    pr.backend = g.DebugBackend(int(np.prod(parts)))
    pr.domain = [0, 4, 0, 2 * np.pi]
    pr.model = g.CartesianDiscreteModel(pr.backend, parts, pr.domain, cells, isperiodic=(False, True))
    pr.reffe = g.ReferenceFE("lagrangian", float, 2)
    pr.V = g.TestFESpace(pr.model, pr.reffe, dirichlet_tags="boundary")
    pr.U = g.TrialFESpace(u, pr.V)
    pr.strategy = strategy
    pr.trian = g.Triangulation(g.FullyAssembledRows(), pr.model) if strategy == "fully" else g.Triangulation(pr.model)
    pr.D, pr.order, pr.ncomp = 2, 2, 1
    return pr, u


@pytest.mark.parametrize("parts", [(1, 1), (2, 2), (1, 4), (2, 3)])
def test_periodic_poisson_solution_kat(parts):
    """reference test/PeriodicBCsTests.jl:8-37 (part grids of test/sequential/PeriodicBCsTests.jl:25-52): domain (0,4)x(0,2π),
    20x20 cells, periodic in y, u = sin(y+π/6) x, f = -Δu, Q2, Dirichlet on the (remaining) boundary:
    sqrt(sum(∫ abs2(u - uh))) < 0.00122."""
    pr, u = _periodic_problem(parts)
    f = lambda x: np.sin(x[1] + np.pi / 6) * x[0]          # -Δu
    out, _ = oracle_assemble(pr, ("poisson",), source=f)
    A, b = gather_global(out)
    x = spla.spsolve(A.tocsc(), b)
    err = l2_error(pr, x, u)
    assert err < 0.00122
    assert err > 1e-6   # a discretisation error, not an interpolation identity


def test_facet_integration_kats():
    """Boundary facet terms (reference test/PoissonTests.jl:22-39): the oracle's facet quadrature against closed forms and against
    the independent 2-D helper used since round 1."""
    from helpers import oracle_facet_cellvecs, orc
    # surface measure and outward normals of the six faces of a sheared box
    A = np.array([[2.0, 0.3, 0.0], [0.0, 1.5, 0.2], [0.0, 0.0, 1.0]])
    X = np.array([[[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)]], float) @ A.T
    area = {0: np.linalg.norm(np.cross(A[:, 0], A[:, 1])), 2: np.linalg.norm(np.cross(A[:, 0], A[:, 2])), 4: np.linalg.norm(np.cross(A[:, 1], A[:, 2]))}
    for lf in range(6):
        xi, xq, ds, nrm, w = orc.facet_quadrature(X, [lf], 4)
        assert abs((ds[0] * w).sum() - area[lf - lf % 2]) < 1e-12
        inward = X[0].mean(0) - xq[0].mean(0)
        assert (nrm[0] @ inward < 0).all() and np.allclose(np.linalg.norm(nrm[0], axis=1), 1.0)
    # sum_i F_i = ∫_Γ g (partition of unity), g = x*y on the top face of the unit cube, Q2
    X1 = np.array([[[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1)]], float)
    xi, xq, ds, nrm, w = orc.facet_quadrature(X1, [1], 4)
    ref = g.NCube(3).q2_ref_nodes()
    F = orc.integrate_boundary_facets(X1, [1], ref, 2, 1, 4, (xq[..., 0] * xq[..., 1]))
    assert abs(F.sum() - 0.25) < 1e-13
    # the reference configuration: same cell vectors as the round-1 helper
    u = lambda x: (x[0] + x[1]) ** 2
    pr = build_problem((2, 2), (4, 4), 2, [1, 2, 3, 5, 7], u, "sub", domain=[0, 4, 0, 4])
    gN = lambda x, n: 2 * (x[0] + x[1]) * (n[0] + n[1])
    a = neumann_cellvec_2d(pr, gN)
    b = oracle_facet_cellvecs(pr, g.Boundary(pr.model, tags=[4, 6, 8]), gN, 4)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=1e-13, atol=1e-13)


def test_boundary_triangulation_and_facet_points_of_the_host_mirror():
    """Host mirror of ``Boundary(model,tags)`` (reference Geometry.jl:684-767): facet counts per tag on a partitioned model, and its
    facet quadrature points / outward normals (what the caller evaluates g on) against the oracle's, on perturbed hexes."""
    from helpers import orc
    pr = build_problem((2, 1, 2), (4, 3, 4), 2, None, None, "sub")
    G = g.Boundary(pr.model, tags="boundary")
    own = [set(ids.own_to_local.tolist()) for ids in pr.model.cell_gids.indices]
    n_owned_facets = sum(sum(1 for c in cells if int(c) in o) for cells, o in zip(G.cell_lids, own))
    assert n_owned_facets == 2 * (4 * 3 + 3 * 4 + 4 * 4)            # every boundary facet of the 4 x 3 x 4 box exactly once among the owned cells
    top = g.Boundary(pr.model, tags=[22])                            # the z = 1 face only (HEX faces: 21 z=0, 22 z=1, 23 y=0, 24 y=1, 25 x=0, 26 x=1)
    assert all((lf == 1).all() for lf in top.lfaces) and sum(sum(1 for c in cells if int(c) in o) for cells, o in zip(top.cell_lids, own)) == 12
    per = g.CartesianDiscreteModel(g.DebugBackend(2), (2, 1), [0, 1, 0, 1], (4, 4), isperiodic=(True, False))
    Gp = g.Boundary(per, tags="boundary")                            # a periodic direction has no boundary
    assert all(((lf // 2) == 0).all() for lf in Gp.lfaces)           # QUAD facets 0, 1 = y faces only
    pert = g.vertex_perturbation(0.2, seed=4)
    for m, cells, lfaces in zip(pr.model.models, G.cell_lids, G.lfaces):
        X = pert(m, m.vertex_coordinates())[m.cell_vertex_ids()[cells - 1] - 1]
        xq, nrm = g.facet_points(X, lfaces, 4)
        xi, xq2, ds, n2, w = orc.facet_quadrature(X, lfaces, 4)
        assert np.allclose(xq, xq2, rtol=0, atol=1e-14) and np.allclose(nrm, n2, rtol=0, atol=1e-13)


@pytest.mark.parametrize("parts,cells,order", [((2, 1, 1), (4, 3, 3), 2), ((2, 2), (5, 4), 2), ((1, 2, 1), (3, 4, 3), 1)])
def test_patch_test_on_general_cells(parts, cells, order):
    """General (tri)linear cells -- per-quadrature-point Jacobians, the path the sum-factorised and tensor-core kernels are compared
    with -- pinned by properties that do not depend on the oracle itself: (1) constants are in the kernel of every cell matrix and
    the cell volumes add up to the domain; (2) the patch test: with interior vertices moved by up to 0.2 h, a LINEAR u is reproduced
    by the isoparametric discretisation to round-off for any mesh (so the discrete solve of the assembled system returns it)."""
    D = len(cells)
    u = lambda x: 1.0 + sum((d + 1.5) * x[d] for d in range(D))
    pr = build_problem(parts, cells, order, "boundary", u, "sub")
    pert = g.vertex_perturbation(0.2, seed=11)
    coords = lambda m, lids, X: pert(m, m.vertex_coordinates())[m.cell_vertex_ids()[lids - 1] - 1]
    out, KF = oracle_assemble(pr, ("poisson",), source=0.0, perturb=coords)
    vol = 0.0
    for k, (m, s) in enumerate(zip(pr.model.models, pr.U.spaces)):
        lids = pr.trian.cell_lids[k]
        X = coords(m, lids, None)
        K, _ = orc.integrate_cells(("poisson",), X, s.ref_nodes, order, 1, 2 * order)
        assert np.abs(K.sum(2)).max() < 1e-11 and np.abs(K - K.transpose(0, 2, 1)).max() < 1e-12
        M, _ = orc.integrate_cells(("mass",), X, s.ref_nodes, order, 1, 2 * order)
        vol += M.sum()
        assert not np.allclose(X, cell_coords_of(m, lids))          # the mesh really is perturbed
    assert abs(vol - 1.0) < 1e-12                                   # boundary vertices stay: the cells still tile the unit box
    A, b = gather_global(out)
    x = spla.spsolve(A.tocsc(), b)
    # nodal values of u at the free dofs of the PERTURBED mesh: isoparametric nodes follow the geometry map
    err = 0.0
    for k, (m, s, ids) in enumerate(zip(pr.model.models, pr.U.spaces, pr.U.gids.indices)):
        lids = pr.trian.cell_lids[k]
        X = coords(m, lids, None)
        N, _ = orc.geometry_tables(D, s.ref_nodes)
        xn = np.einsum("iv,cvd->cid", N, X)                          # physical position of every local dof node
        cd = s.cell_dof_ids[lids - 1]
        free = cd > 0
        gid = ids.l2g[cd[free] - 1] - 1
        err = max(err, np.abs(x[gid] - u(xn[free].T)).max())
    assert err < 1e-10


def cell_coords_of(m, lids):
    return m.vertex_coordinates()[m.cell_vertex_ids()[lids - 1] - 1]


def test_rigid_body_modes_of_elasticity_on_general_cells():
    """The elasticity cell matrices of perturbed hexes annihilate the six rigid-body modes (three translations, three infinitesimal
    rotations) -- a property of eps(u), independent of the oracle's own tables -- and are symmetric."""
    pr = build_problem((1, 1, 1), (2, 2, 2), 2, None, None, "sub", ncomp=3)
    m, s = pr.model.models[0], pr.U.spaces[0]
    pert = g.vertex_perturbation(0.2, seed=2)
    lids = pr.trian.cell_lids[0]
    X = pert(m, m.vertex_coordinates())[m.cell_vertex_ids()[lids - 1] - 1]
    K, _ = orc.integrate_cells(("elasticity", 1.7, 0.6), X, s.ref_nodes, 2, 3, 4)
    N, _ = orc.geometry_tables(3, s.ref_nodes)
    xn = np.einsum("iv,cvd->cid", N, X)                              # (ncells, 27, 3)
    nds = xn.shape[1]
    modes = []
    for c in range(3):
        t = np.zeros((len(X), 3, nds)); t[:, c] = 1.0
        modes.append(t)
    for a, b in ((0, 1), (1, 2), (0, 2)):
        r = np.zeros((len(X), 3, nds)); r[:, a] = -xn[:, :, b]; r[:, b] = xn[:, :, a]
        modes.append(r)
    scale = np.abs(K).max()
    for v in modes:
        ldof = v.reshape(len(X), 3 * nds)                            # local dof = node + nnodes*comp
        assert np.abs(np.einsum("cij,cj->ci", K, ldof)).max() < 1e-11 * scale * max(1.0, np.abs(ldof).max())
    assert np.abs(K - K.transpose(0, 2, 1)).max() < 1e-12 * scale
