"""Shared test drivers: build a distributed problem with the host mirror, run the CPU oracle on it."""
import numpy as np
import scipy.sparse as sp

import graft_import
from oracle import assembly_oracle as orc

g = graft_import.load()


Problem = g.problems.Problem
build_problem = g.build_problem   # the input producers live in the package (bench.py uses them without importing the oracle)


def cell_coords(model_local, cell_lids):
    xyz = model_local.vertex_coordinates()
    cv = model_local.cell_vertex_ids()[cell_lids - 1]
    return xyz[cv - 1]


def oracle_dofs(pr):
    return [orc.local_indices(i.n_global, i.part, i.l2g, i.l2o) for i in pr.U.gids.indices]


def cell_values(pr, k, free_values):
    """Values of an FE function on the dofs of the integrated cells of part k (free dofs: free_values, Dirichlet dofs: the
    Dirichlet values of the trial space)."""
    s = pr.U.spaces[k]
    ids = s.cell_dof_ids[pr.trian.cell_lids[k] - 1]
    dv = np.asarray(pr.U.dirichlet_values[k], dtype=np.float64)
    d = dv if len(dv) else np.zeros(1)
    fv = np.asarray(free_values, dtype=np.float64)
    fv = fv if len(fv) else np.zeros(1)
    return np.where(ids > 0, fv[np.maximum(ids, 1) - 1], d[np.clip(-ids, 1, len(d)) - 1])


def oracle_assemble(pr, form=("poisson",), source=None, quad_degree=None, extra_cellvec=None, perturb=None, state=None):
    """Run the oracle on every part.  Returns (per-part results, per-part (K,F) cell arrays).
    form = ("plaplacian",) with state = per-part free values of uh: jacobian and residual at uh (no lifting)."""
    qd = quad_degree or 2 * pr.order
    I, J, V, B, T, KF = [], [], [], [], [], []
    dofs = oracle_dofs(pr)
    for k, (m, s) in enumerate(zip(pr.model.models, pr.U.spaces)):
        lids = pr.trian.cell_lids[k]
        X = cell_coords(m, lids)
        if perturb is not None:
            X = perturb(m, lids, X)
        src = source[1][k] if (isinstance(source, tuple) and source[0] == "nodal") else source
        fk = ("plaplacian", cell_values(pr, k, state[k])) if form[0] == "plaplacian" else form
        K, F = orc.integrate_cells(fk, X, s.ref_nodes, pr.order, pr.ncomp, qd, src)
        ids = s.cell_dof_ids[lids - 1]
        if form[0] != "plaplacian":   # the increment of a nonlinear problem vanishes on the Dirichlet boundary: nothing to lift
            F = orc.lift_dirichlet(K, F, ids, pr.U.dirichlet_values[k])
        if extra_cellvec is not None:
            F = F + extra_cellvec[k]
        mask = (dofs[k]["l2o"] != dofs[k]["part"]) if pr.strategy == "fully" else None
        i, j, v, b, t = orc.numeric_loop(ids, ids, K, F, s.num_free_dofs, mask)
        I.append(i); J.append(j); V.append(v); B.append(b); T.append(t); KF.append((K, F))
    out = orc.create_from_nz(pr.strategy, I, J, V, B, T, dofs, dofs)
    return out, KF


def gather_global(out):
    """Global scipy CSR + rhs from the owned rows of every part (what `\\` gathers, README.md:76)."""
    n = out[0]["rows"]["n"]
    rows, cols, vals = [], [], []
    b = np.zeros(n)
    for p in out:
        rowptr, colind, val = p["csr"]
        nown = len(p["rows"]["own_to_local"])
        rid = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
        keep = rid < nown
        rows.append(p["rows"]["l2g"][rid[keep]] - 1)
        cols.append(p["cols"]["l2g"][colind[keep]] - 1)
        vals.append(val[keep])
        if p["b"] is not None:
            b[p["rows"]["l2g"][:nown] - 1] += p["b"][:nown]
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return A, b


def l2_error(pr, xglob, ufun, quad_degree=None):
    """sqrt(sum ∫ |u-uh|^2 dΩ) over owned cells (reference test/PoissonTests.jl:43-45)."""
    qd = quad_degree or 2 * pr.order
    tot = 0.0
    for k, (m, s, ids) in enumerate(zip(pr.model.models, pr.U.spaces, pr.U.gids.indices)):
        own = pr.model.cell_gids.indices[k].own_to_local
        X = cell_coords(m, own)
        phi, dphi, w, xi = orc.reference_tables(pr.D, pr.order, s.ref_nodes, qd)
        N, dN = orc.geometry_tables(pr.D, xi)
        cd = s.cell_dof_ids[own - 1]
        free = xglob[ids.l2g[np.maximum(cd, 1) - 1] - 1]
        dv = pr.U.dirichlet_values[k]
        dirv = dv[np.clip(-cd, 1, max(len(dv), 1)) - 1] if len(dv) else np.zeros(cd.shape)
        uc = np.where(cd > 0, free, dirv)
        uh = np.einsum("qi,ci->cq", phi, uc)
        xq = np.einsum("qv,cvd->dcq", N, X)
        ue = ufun(xq.reshape(pr.D, -1)).reshape(uh.shape)
        J = np.einsum("cvd,qva->cqda", X, dN)
        tot += float(np.einsum("q,cq,cq->", w, np.abs(np.linalg.det(J)), (ue - uh) ** 2))
    return np.sqrt(tot)


def neumann_cellvec_2d(pr, gfun, tags=(6, 8), quad_degree=4):
    """∫ v g dΓ on the boundary facets of the integrated cells whose Cartesian entity is in ``tags``
    (2-D: 6 = top, 8 = right; reference test/PoissonTests.jl:19-23,39).  Returned per parent cell."""
    x1, w1 = orc.gauss_legendre_01((quad_degree + 2) // 2)
    out = []
    for k, (m, s) in enumerate(zip(pr.model.models, pr.U.spaces)):
        lids = pr.trian.cell_lids[k]
        X = cell_coords(m, lids)
        ci = m.cell_multi_index()[lids - 1] + m.cmin[None, :]
        F = np.zeros((len(lids), s.nd))
        for tag, (axis, side) in {5: (1, 0), 6: (1, 1), 7: (0, 0), 8: (0, 1)}.items():
            if tag not in tags:
                continue
            sel = np.flatnonzero(ci[:, axis] == (m.ncells_global[axis] - 1 if side else 0))
            if len(sel) == 0:
                continue
            other = 1 - axis
            xi = np.zeros((len(x1), 2)); xi[:, axis] = side; xi[:, other] = x1
            val = [orc.lagrange_1d(pr.order, xi[:, d])[0] for d in range(2)]
            tix = np.rint(s.ref_nodes * pr.order).astype(int)
            phi = val[0][:, tix[:, 0]] * val[1][:, tix[:, 1]]
            N, _ = orc.geometry_tables(2, xi)
            xq = np.einsum("qv,cvd->dcq", N, X[sel])
            length = np.abs(X[sel][:, -1, other] - X[sel][:, 0, other])
            normal = np.zeros(2); normal[axis] = 1.0 if side else -1.0
            gq = gfun(xq.reshape(2, -1), normal).reshape(len(sel), -1)
            F[sel] += np.einsum("q,c,qi,cq->ci", w1, length, phi, gq)
        out.append(F)
    return out


def stokes_problem(parts, cells, strategy, ufun=None, domain=None):
    pr = g.build_stokes_problem(parts, cells, strategy, ufun=ufun) if domain is None else _stokes_problem_on(parts, cells, strategy, ufun, domain)
    return pr


def _stokes_problem_on(parts, cells, strategy, ufun, domain):
    D = len(cells)
    pr = Problem()
    pr.backend = g.DebugBackend(int(np.prod(parts)))
    pr.model = g.CartesianDiscreteModel(pr.backend, parts, domain, cells)
    pr.V = g.TestFESpace(pr.model, g.ReferenceFE("lagrangian", float, 2, ncomp=D), dirichlet_tags="boundary")
    pr.Q = g.TestFESpace(pr.model, g.ReferenceFE("lagrangian", float, 1), dirichlet_tags=None)
    pr.U = g.TrialFESpace(ufun, pr.V)
    pr.P = g.TrialFESpace(None, pr.Q)
    pr.strategy, pr.D = strategy, D
    pr.trian = g.Triangulation(g.FullyAssembledRows(), pr.model) if strategy == "fully" else g.Triangulation(pr.model)
    return pr


def stokes_oracle(pr, nu, source, perturb=None, extra_p=None):
    """The oracle's block pipeline on the Taylor-Hood Stokes forms; extra_p: per-part additive pressure cell vectors."""
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
        if perturb is not None:
            X = perturb(m, lids, X)
        su, sp_ = pr.U.spaces[k], pr.P.spaces[k]
        Kuu, Kup, Kpu, Fu = orc.integrate_stokes_cells(X, su.ref_nodes, sp_.ref_nodes, 2, 1, 4, nu, source)
        idu, idp = su.cell_dof_ids[lids - 1], sp_.cell_dof_ids[lids - 1]
        Fp = np.zeros((len(lids), idp.shape[1])) if extra_p is None else np.array(extra_p[k], dtype=np.float64)
        # lifting with the Dirichlet values of both trial fields (FESpaces.jl:703-715; the pressure has none here)
        Fu = orc.lift_dirichlet(Kuu, Fu, idu, pr.U.dirichlet_values[k])
        Fu = orc.lift_dirichlet(Kup, Fu, idp, pr.P.dirichlet_values[k])
        Fp = orc.lift_dirichlet(Kpu, Fp, idu, pr.U.dirichlet_values[k])
        masks = [(d[k]["l2o"] != d[k]["part"]) if pr.strategy == "fully" else None for d in dofs]
        I[0][0][k], J[0][0][k], V[0][0][k], B[0][k], T[0][k] = orc.numeric_loop(idu, idu, Kuu, Fu, su.num_free_dofs, masks[0])
        I[0][1][k], J[0][1][k], V[0][1][k], _, _ = orc.numeric_loop(idu, idp, Kup, None, su.num_free_dofs, masks[0])
        I[1][0][k], J[1][0][k], V[1][0][k], B[1][k], T[1][k] = orc.numeric_loop(idp, idu, Kpu, Fp, sp_.num_free_dofs, masks[1])
    I[1][1] = J[1][1] = V[1][1] = None  # the form does not touch the (p,p) block
    return orc.create_from_nz_blocks(pr.strategy, I, J, V, B, T, dofs, dofs)




def block_oracle_from_cell_arrays(pr, spaces, mats, vecs):
    """The oracle's block pipeline (numeric_loop per block + create_from_nz_blocks) on caller-supplied cell arrays:
    mats[(i,j)][part] (ncells, nd_i, nd_j) -- a missing key is a block the form does not touch --, vecs[i][part]."""
    nf = len(spaces)
    dofs = [[orc.local_indices(i.n_global, i.part, i.l2g, i.l2o) for i in sp.gids.indices] for sp in spaces]
    P = len(pr.model.models)
    I = [[[None] * P if (i, j) in mats else None for j in range(nf)] for i in range(nf)]
    J = [[[None] * P if (i, j) in mats else None for j in range(nf)] for i in range(nf)]
    V = [[[None] * P if (i, j) in mats else None for j in range(nf)] for i in range(nf)]
    B = [[None] * P for _ in range(nf)]
    T = [[None] * P for _ in range(nf)]
    for k in range(P):
        lids = pr.trian.cell_lids[k]
        ids = [sp.spaces[k].cell_dof_ids[lids - 1] for sp in spaces]
        masks = [(d[k]["l2o"] != d[k]["part"]) if pr.strategy == "fully" else None for d in dofs]
        for i in range(nf):
            first = True
            for j in range(nf):
                if (i, j) not in mats:
                    continue
                F = vecs[i][k] if (first and i in vecs) else None
                I[i][j][k], J[i][j][k], V[i][j][k], bb, tt = orc.numeric_loop(ids[i], ids[j], mats[(i, j)][k], F, spaces[i].spaces[k].num_free_dofs,
                                                                               masks[i])
                if first:
                    B[i][k], T[i][k] = bb, tt
                first = False
    if not vecs:
        B = None
    return orc.create_from_nz_blocks(pr.strategy, I, J, V, B, T, dofs, dofs)


def oracle_facet_cellvecs(pr, boundary, gfun, quad_degree, perturb=None, space=None):
    """Per part (n_integrated_cells, nd): the oracle's facet vectors of  ∫ v g dΓ  (oracle.integrate_boundary_facets) added to
    their parent cells -- the CPU counterpart of graft_neumann_set.  boundary: g.Boundary(model, tags)."""
    space = space or pr.U
    out = []
    for k, (m, s) in enumerate(zip(pr.model.models, space.spaces)):
        lids = pr.trian.cell_lids[k]
        X = cell_coords(m, lids)
        if perturb is not None:
            X = perturb(m, lids, X)
        F = np.zeros((len(lids), s.nd))
        pos = np.searchsorted(lids, boundary.cell_lids[k])
        keep = (pos < len(lids)) & (lids[np.minimum(pos, len(lids) - 1)] == boundary.cell_lids[k]) if len(lids) else np.zeros(0, dtype=bool)
        cells, lfaces = pos[keep], boundary.lfaces[k][keep]
        if len(cells):
            xi, xq, ds, nrm, w = orc.facet_quadrature(X[cells], lfaces, quad_degree)
            nf, nqf, D = xq.shape
            vals = np.asarray(gfun(xq.reshape(-1, D).T, nrm.reshape(-1, D).T), dtype=np.float64)
            vals = np.broadcast_to(vals, (s.ncomp, nf * nqf)) if vals.ndim < 2 else vals.reshape(s.ncomp, nf * nqf)
            Ff = orc.integrate_boundary_facets(X[cells], lfaces, s.ref_nodes, pr.order, s.ncomp, quad_degree, vals.T.reshape(nf, nqf, s.ncomp))
            np.add.at(F, cells, Ff)
        out.append(F)
    return out
