"""CPU ORACLE (test infrastructure, NOT product code) for the distributed assembly + mul! path.

A numpy restatement of the reference algorithm, function by function (SURVEY.md section 8a rows
A1-A14).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path never does.

PARITY STATUS -- "parity unpinned at entry level": the reference (GridapDistributed.jl 0.4.x) is
Julia, cannot run in this image, and its own tests hold NO golden matrices, patterns or index maps
(SURVEY.md section 8c).  The arithmetic of the path lives in un-vendored dependencies --
Gridap.jl >=0.20.8,<0.21, PartitionedArrays.jl >=0.3.3,<0.4, SparseMatricesCSR.jl 0.6.6
(reference Project.toml:19-31) -- whose published algorithms are restated here ("[ext]" below).
The oracle is pinned by every known-answer test the reference's tests DO hold for this path
(tests/test_oracle_kats.py): solution error < 1e-9 for u=(x+y)^2 Q2 on 4x4 (2,2)
(test/PoissonTests.jl:45), |sum(b)-length(b)| < 1e-12 (test/FESpacesTests.jl:66), local matrix
size == local lengths (test/FESpacesTests.jl:29-31), FullyAssembledRows == SubAssembledRows
(test/FESpacesTests.jl:73-76), mul! alpha/beta semantics (test/BlockPartitionedArraysTests.jl:70-77),
plus analytic element matrices and nnz closed forms.

All ids are 1-based like the reference; arrays are numpy.  Every part lives in this process
(the reference's ``with_debug`` mode): "exchanges" are list shuffles.
"""
from __future__ import annotations

import numpy as np

# ------------------------------------------------------------------------------------------------
# A1  cell integration  ([ext] Gridap.integrate; called from reference CellData.jl:234-239)
# ------------------------------------------------------------------------------------------------


def gauss_legendre_01(n):
    """n-point Gauss-Legendre rule on [0,1] (Gridap ``Measure(Ω,deg)``: n = ceil((deg+1)/2), A.4)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return (x + 1.0) / 2.0, w / 2.0


def lagrange_1d(order, x):
    """Values and derivatives of the equispaced Lagrange basis of ``order`` on [0,1] at points x.
    Node j sits at j/order.  Returns (val[len(x), order+1], der[len(x), order+1])."""
    nodes = np.arange(order + 1) / order
    x = np.asarray(x, dtype=np.float64)
    val = np.ones((len(x), order + 1))
    der = np.zeros((len(x), order + 1))
    for j in range(order + 1):
        for m in range(order + 1):
            if m != j:
                val[:, j] *= (x - nodes[m]) / (nodes[j] - nodes[m])
        for i in range(order + 1):
            if i == j:
                continue
            t = np.full(len(x), 1.0 / (nodes[j] - nodes[i]))
            for m in range(order + 1):
                if m != j and m != i:
                    t *= (x - nodes[m]) / (nodes[j] - nodes[m])
            der[:, j] += t
    return val, der


def reference_tables(D, order, ref_nodes, quad_degree):
    """phi[q,i], dphi[q,i,a], w[q], xi[q,a] for the scalar tensor Lagrange basis whose local dof
    order is given by ``ref_nodes`` (nd x D reference coordinates: SURVEY A.5, never hard-coded).
    Quadrature points are x fastest (A.4)."""
    n = (quad_degree + 2) // 2  # ceil((deg+1)/2)
    x1, w1 = gauss_legendre_01(n)
    val, der = lagrange_1d(order, x1)
    tix = np.rint(np.asarray(ref_nodes) * order).astype(int)  # tensor index of every local dof
    nd = len(tix)
    nq = n**D
    qidx = np.stack([(np.arange(nq) // n**d) % n for d in range(D)], axis=1)
    xi = x1[qidx]
    w = np.prod(w1[qidx], axis=1)
    phi = np.ones((nq, nd))
    dphi = np.ones((nq, nd, D))
    for d in range(D):
        phi *= val[qidx[:, d]][:, tix[:, d]]
        for a in range(D):
            src = der if a == d else val
            dphi[:, :, a] *= src[qidx[:, d]][:, tix[:, d]]
    return phi, dphi, w, xi


def geometry_tables(D, xi):
    """Q1 geometry map shape functions N[q,v] and gradients dN[q,v,a], vertices x fastest."""
    nv = 2**D
    vc = np.array([[(v >> d) & 1 for d in range(D)] for v in range(nv)], dtype=np.float64)
    N = np.ones((len(xi), nv))
    dN = np.ones((len(xi), nv, D))
    for d in range(D):
        f = np.where(vc[None, :, d] == 1, xi[:, None, d], 1.0 - xi[:, None, d])
        g = np.where(vc[None, :, d] == 1, 1.0, -1.0) * np.ones_like(f)
        N *= f
        for a in range(D):
            dN[:, :, a] *= g if a == d else f
    return N, dN


def facet_quadrature(cell_coords, lfaces, quad_degree):
    """Quadrature on boundary facets (reference ``Boundary(model,tags)`` / ``Measure(Γ,degree)``, Geometry.jl:684-767, and the
    facet maps of Gridap's BoundaryTriangulation [ext, A.4]).  cell_coords: (nf, 2^D, D) vertices of the parent cells; lfaces: (nf,)
    facet of the n-cube in Gridap's order (HEX: z=0, z=1, y=0, y=1, x=0, x=1; QUAD: y=0, y=1, x=0, x=1; SEGMENT: x=0, x=1).
    Returns xi (nf, nqf, D) reference points of the parent cell, xq (nf, nqf, D) physical points, ds (nf, nqf) surface measures,
    normals (nf, nqf, D) outward unit normals, w (nqf,) weights; facet points: tensor Gauss-Legendre over the facet's free axes,
    lowest axis fastest."""
    cell_coords = np.asarray(cell_coords, dtype=np.float64)
    nf, nv, D = cell_coords.shape
    lfaces = np.asarray(lfaces, dtype=np.int64)
    n = (quad_degree + 2) // 2
    x1, w1 = gauss_legendre_01(n)
    nqf = n ** (D - 1)
    qidx = np.stack([(np.arange(nqf) // n**o) % n for o in range(D - 1)], axis=1) if D > 1 else np.zeros((1, 0), dtype=np.int64)
    w = np.prod(w1[qidx], axis=1) if D > 1 else np.ones(1)
    xi = np.zeros((nf, nqf, D))
    axis = D - 1 - lfaces // 2
    side = lfaces % 2
    for f in range(nf):
        others = [d for d in range(D) if d != axis[f]]
        xi[f, :, axis[f]] = side[f]
        for o, d in enumerate(others):
            xi[f, :, d] = x1[qidx[:, o]]
    xq = np.zeros((nf, nqf, D)); ds = np.zeros((nf, nqf)); normals = np.zeros((nf, nqf, D))
    for f in range(nf):
        N, dN = geometry_tables(D, xi[f])
        xq[f] = N @ cell_coords[f]
        J = np.einsum("qva,vd->qda", dN, cell_coords[f])           # J[q][d][a] = dx_d / dxi_a
        for q in range(nqf):
            # Nanson: n ds = det(J) J^{-T} e_axis (reference normal of the facet, outward: -e_axis on the low side)
            v = np.linalg.det(J[q]) * np.linalg.inv(J[q]).T[:, axis[f]] * (1.0 if side[f] else -1.0)
            ds[f, q] = np.linalg.norm(v)
            normals[f, q] = v / ds[f, q]
    return xi, xq, ds, normals, w


def integrate_boundary_facets(cell_coords, lfaces, ref_nodes, order, ncomp, quad_degree, g):
    """Facet vectors of  ∫ v g dΓ  (the term ``∫( v*g )dΓn`` of reference test/PoissonTests.jl:39): F[f, li] = Σ_q w_q ds_q φ_li(ξ_q)
    g[f, q, comp(li)] with the local dof order li = node + nnodes*comp (A.5).  g: (nf, nqf) or (nf, nqf, ncomp)."""
    xi, _, ds, _, w = facet_quadrature(cell_coords, lfaces, quad_degree)
    nf, nqf, D = xi.shape
    tix = np.rint(np.asarray(ref_nodes) * order).astype(int)
    nds = len(tix)
    g = np.asarray(g, dtype=np.float64).reshape(nf, nqf, ncomp)
    F = np.zeros((nf, nds * ncomp))
    for f in range(nf):
        phi = np.ones((nqf, nds))
        for d in range(D):
            val, _ = lagrange_1d(order, xi[f, :, d])
            phi *= val[:, tix[:, d]]
        for c in range(ncomp):
            F[f, c * nds:(c + 1) * nds] = np.einsum("q,q,qi,q->i", w, ds[f], phi, g[f, :, c])
    return F


def integrate_cells(form, cell_coords, ref_nodes, order, ncomp, quad_degree, source=None, chunk=4096):
    """Cell matrices K[c,i,j] (test i = row, trial j = col) and vectors f[c,i] by quadrature (A.4).

    form: ("poisson",) | ("mass",) | ("elasticity", lam, mu).  cell_coords: (ncells, 2^D, D).
    source: None | float | callable f(x[D,npts]) -> [npts] (scalar) or [ncomp,npts] | ndarray (ncells, nd) of
    nodal values of an FE-function source.
    """
    ncells, nv, D = cell_coords.shape
    phi, dphi, w, xi = reference_tables(D, order, ref_nodes, quad_degree)
    N, dN = geometry_tables(D, xi)
    nds = phi.shape[1]
    nd = nds * ncomp
    K = np.zeros((ncells, nd, nd))
    F = np.zeros((ncells, nd))
    for s in range(0, ncells, chunk):
        X = cell_coords[s : s + chunk]
        J = np.einsum("cvd,qva->cqda", X, dN)  # J[d,a] = dx_d/dxi_a
        detJ = np.linalg.det(J)
        Jinv = np.linalg.inv(J)  # [a,d]
        g = np.einsum("qia,cqad->cqid", dphi, Jinv)  # physical gradients
        wd = w[None, :] * np.abs(detJ)
        if form[0] == "poisson":
            Ks = np.einsum("cq,cqid,cqjd->cij", wd, g, g)
        elif form[0] == "mass":
            Ks = np.einsum("cq,qi,qj->cij", wd, phi, phi)
        elif form[0] == "elasticity":
            lam, mu = form[1], form[2]
            Kab = np.einsum("cq,cqia,cqjb->ciajb", wd, g, g)  # int d_a phi_i d_b phi_j
            Ks = None
            Kv = np.zeros((len(X), ncomp, nds, ncomp, nds))
            lap = np.einsum("ciaja->cij", Kab)
            for al in range(ncomp):
                for be in range(ncomp):
                    blk = lam * Kab[:, :, al, :, be] + mu * Kab[:, :, be, :, al]
                    if al == be:
                        blk = blk + mu * lap
                    Kv[:, al, :, be, :] = blk
            K[s : s + chunk] = Kv.reshape(len(X), nd, nd)
        elif form[0] == "plaplacian":
            # state-dependent forms of reference test/PLaplacianTests.jl:24-31 at the state uh (nodal cell values form[1]):
            #   sigma(grad u) = (1 + grad u . grad u) grad u ,  dsigma(grad du, grad u) = 2 (grad u . grad du) grad u + (1 + |grad u|^2) grad du
            #   jacobian  K[i,j] = int grad phi_i . dsigma(grad phi_j, grad uh) ,  residual  F[i] = int grad phi_i . sigma(grad uh) - phi_i f
            uc = np.asarray(form[1])[s : s + chunk]                      # (ncells, nds)
            gu = np.einsum("cqjd,cj->cqd", g, uc)                        # grad uh at the points
            n2 = np.einsum("cqd,cqd->cq", gu, gu)
            gi_gu = np.einsum("cqid,cqd->cqi", g, gu)
            Ks = np.einsum("cq,cqi,cqj->cij", 2.0 * wd, gi_gu, gi_gu) + np.einsum("cq,cqid,cqjd->cij", wd * (1.0 + n2), g, g)
            F[s : s + chunk] = np.einsum("cq,cqi->ci", wd * (1.0 + n2), gi_gu)
        else:
            raise ValueError(form)
        if form[0] != "elasticity":
            for c in range(ncomp):  # componentwise scalar forms
                K[s : s + chunk, c * nds : (c + 1) * nds, c * nds : (c + 1) * nds] = Ks
        if source is not None:
            if isinstance(source, np.ndarray):
                # FE-function source: nodal values per cell dof (ncells, nd); f(x_q) = sum_k phi_k(xi_q) F_k
                Fn = source[s : s + chunk].reshape(len(X), ncomp, nds)
                fq = np.einsum("qk,cak->acq", phi, Fn)
            elif callable(source):
                xq = np.einsum("qv,cvd->dcq", N, X).reshape(D, -1)
                fq = np.asarray(source(xq), dtype=np.float64)
                fq = fq.reshape(ncomp, len(X), -1) if fq.ndim == 2 else np.broadcast_to(fq.reshape(1, len(X), -1), (ncomp, len(X), len(w)))
            else:
                fq = np.full((ncomp, len(X), len(w)), float(source))
            for c in range(ncomp):
                if form[0] == "plaplacian":   # residual: ... - int v f
                    F[s : s + chunk, c * nds : (c + 1) * nds] -= np.einsum("cq,qi,cq->ci", wd, phi, fq[c])
                else:
                    F[s : s + chunk, c * nds : (c + 1) * nds] = np.einsum("cq,qi,cq->ci", wd, phi, fq[c])
    return K, F


# ------------------------------------------------------------------------------------------------
# A2  Dirichlet lifting  ([ext] collect_cell_matrix_and_vector(U,V,a,l,uhd); reference FESpaces.jl:703-715)
# ------------------------------------------------------------------------------------------------


def lift_dirichlet(K, F, cell_dof_ids, dirichlet_values):
    """f_c <- f_c - K_c u_c with u_c = Dirichlet values on Dirichlet dofs, 0 on free dofs (A.7)."""
    ids = cell_dof_ids
    dv = np.asarray(dirichlet_values, dtype=np.float64)
    u = np.where(ids < 0, dv[np.clip(-ids, 1, max(len(dv), 1)) - 1], 0.0) if len(dv) else np.zeros(ids.shape)
    return F - np.einsum("cij,cj->ci", K, u)


# ------------------------------------------------------------------------------------------------
# A3  numeric loop  ([ext] add_entries!; reference FESpaces.jl:794-798, strategies :802-816,
#      vector allocation Algebra.jl:808-821)
# ------------------------------------------------------------------------------------------------


def numeric_loop(cell_dof_ids_rows, cell_dof_ids_cols, K, F, nrows_local, row_is_ghost=None):
    """Triplets in the reference's order: per cell, lj outer, li inner, ids<=0 skipped (A.3).
    ``row_is_ghost`` (bool per FE-space local row) is the FullyAssembledRows mask
    (``rows_local_to_ghost[row]==0``, FESpaces.jl:808-816); it masks vector rows too."""
    rid, cid = cell_dof_ids_rows, cell_dof_ids_cols
    rvalid = rid > 0
    if row_is_ghost is not None:
        rvalid &= ~row_is_ghost[np.maximum(rid, 1) - 1]
    cvalid = cid > 0
    ncells, ndr = rid.shape
    ndc = cid.shape[1]
    # order (cell, lj, li): build [c, lj, li] arrays
    m = cvalid[:, :, None] & rvalid[:, None, :]
    I = np.broadcast_to(rid[:, None, :], m.shape)[m].astype(np.int64)
    J = np.broadcast_to(cid[:, :, None], m.shape)[m].astype(np.int64)
    V = np.transpose(K, (0, 2, 1))[m]  # K[c, li, lj] viewed as [c, lj, li]
    b = np.zeros(nrows_local)
    touched = np.zeros(nrows_local, dtype=bool)
    if F is not None:
        # b[i] = b[i] + f[li] sequentially over cells (Algebra.jl:811-821): np.add.at is sequential
        np.add.at(b, rid[rvalid] - 1, F[rvalid])
        touched[rid[rvalid] - 1] = True
    return I, J, V, b, touched


# ------------------------------------------------------------------------------------------------
# index-set helpers (reference Algebra.jl:919-1009, 1174-1257); an index set is a dict
# ------------------------------------------------------------------------------------------------


def local_indices(n_global, part, l2g, l2o):
    l2g = np.asarray(l2g, dtype=np.int64)
    l2o = np.asarray(l2o, dtype=np.int32)
    own = l2o == part
    o2l = np.flatnonzero(own) + 1
    g2l_ = np.flatnonzero(~own) + 1
    l2gh = np.zeros(len(l2g), dtype=np.int64)
    l2gh[g2l_ - 1] = np.arange(1, len(g2l_) + 1)
    return dict(n=int(n_global), part=int(part), l2g=l2g, l2o=l2o, own_to_local=o2l, ghost_to_local=g2l_,
                local_to_ghost=l2gh, lookup=dict(zip(l2g.tolist(), range(1, len(l2g) + 1))))


def own_and_ghost(n_global, part, own_to_global, ghost_to_global, ghost_to_owner):
    l2g = np.concatenate([np.asarray(own_to_global, np.int64), np.asarray(ghost_to_global, np.int64)])
    l2o = np.concatenate([np.full(len(own_to_global), part, np.int32), np.asarray(ghost_to_owner, np.int32)])
    d = local_indices(n_global, part, l2g, l2o)
    d.update(own_to_global=np.asarray(own_to_global, np.int64), ghost_to_global=np.asarray(ghost_to_global, np.int64),
             ghost_to_owner=np.asarray(ghost_to_owner, np.int32))
    return d


def g2l(ids, gids):
    lk = ids["lookup"]
    return np.fromiter((lk.get(int(g), 0) for g in gids), dtype=np.int64, count=len(gids))


def first_appearance_unique(a):
    u, first = np.unique(a, return_index=True)
    return u[np.argsort(first, kind="stable")]


def ghost_lids_touched(ids, gids):  # Algebra.jl:931-962
    lid = g2l(ids, gids)
    gh = ids["local_to_ghost"][lid - 1]
    return first_appearance_unique(gh[gh > 0])


def find_gid_and_owner(ighost_to_jghost, jids):  # Algebra.jl:919-928 (note the sort at :923)
    jl = np.sort(jids["ghost_to_local"][np.asarray(ighost_to_jghost, dtype=np.int64) - 1])
    return jids["l2g"][jl - 1], jids["l2o"][jl - 1]


def setup_prange_without_ghosts(dofs):  # Algebra.jl:1174-1183
    return own_and_ghost(dofs["n"], dofs["part"], dofs["l2g"][dofs["own_to_local"] - 1], [], [])


def setup_prange_with_ghosts(dofs, gids, owners=None):
    own = dofs["l2g"][dofs["own_to_local"] - 1]
    if owners is None:  # Algebra.jl:1191-1203
        g, o = find_gid_and_owner(ghost_lids_touched(dofs, gids), dofs)
    else:  # Algebra.jl:1210-1238: first appearance among entries whose column owner is not me
        sel = owners != dofs["part"]
        gs, os_ = gids[sel], owners[sel]
        u, first = np.unique(gs, return_index=True)
        order = np.argsort(first, kind="stable")
        g, o = u[order], os_[first[order]]
    return own_and_ghost(dofs["n"], dofs["part"], own, g, o)


def assembly_neighbors(rows_all):
    """[ext] parts_snd = sorted owners of my ghosts; parts_rcv = ascending parts that send to me."""
    snd = [sorted(set(r["ghost_to_owner"].tolist())) for r in rows_all]
    rcv = [[q + 1 for q, s in enumerate(snd) if (p + 1) in s] for p in range(len(rows_all))]
    return snd, rcv


# ------------------------------------------------------------------------------------------------
# A6  ghost-row triplet migration (reference Algebra.jl:1048-1131)
# ------------------------------------------------------------------------------------------------


def assemble_coo_with_column_owner(I, J, V, rows_all, Jo):
    P = len(I)
    parts_snd, parts_rcv = assembly_neighbors(rows_all)
    outbox = [dict() for _ in range(P)]
    for p in range(P):
        rows = rows_all[p]
        li = g2l(rows, I[p])
        owner = rows["l2o"][li - 1]
        for q in parts_snd[p]:
            k = np.flatnonzero(owner == q)  # in triplet order (:1072-1087)
            outbox[p][q] = (I[p][k].copy(), J[p][k].copy(), Jo[p][k].copy(), V[p][k].copy())
            V[p][k] = 0.0  # :1084 -- I,J stay, values zeroed
    for p in range(P):
        for q in parts_rcv[p]:  # appended in neighbour order (:1099-1112)
            gi, gj, jo, v = outbox[q - 1][p + 1]
            I[p] = np.concatenate([I[p], gi])
            J[p] = np.concatenate([J[p], gj])
            Jo[p] = np.concatenate([Jo[p], jo])
            V[p] = np.concatenate([V[p], v])
    return I, J, Jo, V


# ------------------------------------------------------------------------------------------------
# A9  COO -> CSR  ([ext] sparsecsr / sparse: stable sort, duplicates summed in sorted order)
# ------------------------------------------------------------------------------------------------


def coo_to_csr(I, J, V, m, n):
    """Returns (rowptr[m+1], colind[nnz], nzval[nnz]) 0-based; explicit zeros are kept."""
    I = np.asarray(I, np.int64) - 1
    J = np.asarray(J, np.int64) - 1
    order = np.lexsort((J, I))  # stable: equal (i,j) keep triplet order
    Is, Js, Vs = I[order], J[order], np.asarray(V)[order]
    if len(Is) == 0:
        return np.zeros(m + 1, np.int64), np.zeros(0, np.int64), np.zeros(0)
    new = np.ones(len(Is), dtype=bool)
    new[1:] = (Is[1:] != Is[:-1]) | (Js[1:] != Js[:-1])
    seg = np.cumsum(new) - 1
    start = np.flatnonzero(new)
    rank = np.arange(len(Is)) - start[seg]
    vals = np.zeros(len(start))
    for r in range(int(rank.max()) + 1):  # sequential left-to-right sum per duplicate group
        s = rank == r
        vals[seg[s]] += Vs[s]
    rows = Is[start]
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    return np.cumsum(rowptr), Js[start], vals


# ------------------------------------------------------------------------------------------------
# A11  right-hand side (reference Algebra.jl:720-753, 870-907) and PVector assemble! ([ext] A.9)
# ------------------------------------------------------------------------------------------------


def prange_from_touched(dofs, touched):  # Algebra.jl:870-896
    lids = np.flatnonzero(touched) + 1
    gh = dofs["local_to_ghost"][lids - 1]
    g, o = find_gid_and_owner(gh[gh > 0], dofs)
    return own_and_ghost(dofs["n"], dofs["part"], dofs["l2g"][dofs["own_to_local"] - 1], g, o)


def rhs_callback(b_fespace, dofs, rows):  # Algebra.jl:720-753
    b = np.zeros(len(rows["l2g"]))
    nown = len(rows["own_to_local"])
    b[:nown] = b_fespace[dofs["own_to_local"] - 1]  # :730 own values (same own order)
    gl = g2l(dofs, rows["l2g"][nown:])  # :733-750
    b[nown:] = b_fespace[gl - 1]
    return b


def assemble_pvector(b_all, rows_all):
    """``assemble!(b)``: ghost -> owner with +, in neighbour order, then ghosts are zeroed (A.9)."""
    P = len(b_all)
    parts_snd, parts_rcv = assembly_neighbors(rows_all)
    box = [dict() for _ in range(P)]
    for p in range(P):
        r = rows_all[p]
        nown = len(r["own_to_local"])
        for q in parts_snd[p]:
            k = np.flatnonzero(r["ghost_to_owner"] == q)
            box[p][q] = (r["ghost_to_global"][k], b_all[p][nown + k].copy())
        b_all[p][nown:] = 0.0
    for p in range(P):
        for q in parts_rcv[p]:
            g, v = box[q - 1][p + 1]
            np.add.at(b_all[p], g2l(rows_all[p], g) - 1, v)
    return b_all


# ------------------------------------------------------------------------------------------------
# A12  orchestration (reference Algebra.jl:593-646 SubAssembledRows, :549-579 FullyAssembledRows)
# ------------------------------------------------------------------------------------------------


def create_from_nz(strategy, I, J, V, b_fe, touched, test_dofs, trial_dofs):
    """Per-part lists in, per-part (rows, cols, (rowptr,colind,nzval), b) out."""
    P = len(I)
    I = [np.array(x, np.int64) for x in I]
    J = [np.array(x, np.int64) for x in J]
    V = [np.array(x, np.float64) for x in V]
    if strategy == "fully":
        rows = [setup_prange_without_ghosts(d) for d in test_dofs]  # :556
        b = [rhs_callback(b_fe[p], test_dofs[p], rows[p]) for p in range(P)] if b_fe is not None else None
        I = [test_dofs[p]["l2g"][I[p] - 1] for p in range(P)]  # :560-561
        J = [trial_dofs[p]["l2g"][J[p] - 1] for p in range(P)]
        cols = [setup_prange_with_ghosts(trial_dofs[p], J[p]) for p in range(P)]  # :564
    else:
        I = [test_dofs[p]["l2g"][I[p] - 1] for p in range(P)]  # :600-601
        J = [trial_dofs[p]["l2g"][J[p] - 1] for p in range(P)]
        rows = [setup_prange_with_ghosts(test_dofs[p], I[p]) for p in range(P)]  # :604
        Jo = [trial_dofs[p]["l2o"][g2l(trial_dofs[p], J[p]) - 1] for p in range(P)]  # :608
        I, J, Jo, V = assemble_coo_with_column_owner(I, J, V, rows, Jo)  # :609
        b = None
        if b_fe is not None:
            brows = [prange_from_touched(test_dofs[p], touched[p]) for p in range(P)]  # :615
            b = [rhs_callback(b_fe[p], test_dofs[p], brows[p]) for p in range(P)]  # :616
        cols = [setup_prange_with_ghosts(trial_dofs[p], J[p], Jo[p]) for p in range(P)]  # :623
        if b is not None:
            b = assemble_pvector(b, brows)  # :626
    out = []
    for p in range(P):
        li = g2l(rows[p], I[p])  # :629-630 / :567-568
        lj = g2l(cols[p], J[p])
        csr = coo_to_csr(li, lj, V[p], len(rows[p]["l2g"]), len(cols[p]["l2g"]))  # :636 / :574
        out.append(dict(rows=rows[p], cols=cols[p], csr=csr, b=None if b is None else b[p],
                        brows=None if (b is None or strategy == "fully") else brows[p]))
    return out



# ------------------------------------------------------------------------------------------------
# A15  block pipeline (reference MultiField.jl:473-560; Algebra.jl:488-496, 508-528, 543-547, 587-591,
#      990-1009, 1017-1024, 1036-1046, 1146-1169)
# ------------------------------------------------------------------------------------------------


def integrate_stokes_cells(cell_coords, ref_u, ref_p, order_u, order_p, quad_degree, nu, source_u=None, chunk=2048):
    """Taylor-Hood Stokes blocks of  a((u,p),(v,q)) = ∫ ν ∇v⊙∇u − (∇⋅v) p − q (∇⋅u) dΩ  (reference
    test/StokesOpenBoundaryTests.jl:44-49) per cell: Kuu[(al,i),(be,j)], Kup[(al,i),j], Kpu[i,(be,j)] and the
    velocity source vector.  Velocity dofs are component-major (ldof = node + nnodes*comp, A.5)."""
    ncells, nv, D = cell_coords.shape
    phiu, dphiu, w, xi = reference_tables(D, order_u, ref_u, quad_degree)
    phip, _, _, _ = reference_tables(D, order_p, ref_p, quad_degree)
    _, dN = geometry_tables(D, xi)
    nu_s, np_s = phiu.shape[1], phip.shape[1]
    Kuu = np.zeros((ncells, D * nu_s, D * nu_s))
    Kup = np.zeros((ncells, D * nu_s, np_s))
    Kpu = np.zeros((ncells, np_s, D * nu_s))
    Fu = np.zeros((ncells, D * nu_s))
    for s in range(0, ncells, chunk):
        X = cell_coords[s : s + chunk]
        J = np.einsum("cvd,qva->cqda", X, dN)
        wd = w[None, :] * np.abs(np.linalg.det(J))
        g = np.einsum("qia,cqad->cqid", dphiu, np.linalg.inv(J))
        lap = nu * np.einsum("cq,cqid,cqjd->cij", wd, g, g)
        gv = np.einsum("cq,cqia,qj->caij", wd, g, phip)  # ∫ d_a phi_i psi_j
        for a in range(D):
            Kuu[s : s + chunk, a * nu_s : (a + 1) * nu_s, a * nu_s : (a + 1) * nu_s] = lap
            Kup[s : s + chunk, a * nu_s : (a + 1) * nu_s, :] = -gv[:, a]
            Kpu[s : s + chunk, :, a * nu_s : (a + 1) * nu_s] = -np.transpose(gv[:, a], (0, 2, 1))
            if source_u is not None:
                Fu[s : s + chunk, a * nu_s : (a + 1) * nu_s] = float(source_u) * np.einsum("cq,qi->ci", wd, phiu)
    return Kuu, Kup, Kpu, Fu


def create_from_nz_blocks(strategy, I, J, V, b_fe, touched, test_dofs, trial_dofs):
    """Block version of create_from_nz.  I[i][j] / J[i][j] / V[i][j]: per-part triplet lists of block (i,j) or None
    (block untouched by the form); b_fe[i], touched[i]: per-part FE-space vectors of field i; test_dofs[i] /
    trial_dofs[j]: per-part index sets.  The row PRange of block-row i is built from the union of I over the blocks
    (i,:), the column PRange of block-column j from the union of J (and owners) over the blocks (:,j), concatenated
    in block order (Algebra.jl:1146-1169); the ghost-row migration runs per block with rows[i] (:1036-1046)."""
    nf = len(test_dofs)
    P = len(test_dofs[0])
    act = [[I[i][j] is not None for j in range(nf)] for i in range(nf)]
    G = lambda dofs, L: [dofs[p]["l2g"][np.asarray(L[p], np.int64) - 1] for p in range(P)]
    Ig = [[G(test_dofs[i], I[i][j]) if act[i][j] else None for j in range(nf)] for i in range(nf)]  # :600-601
    Jg = [[G(trial_dofs[j], J[i][j]) if act[i][j] else None for j in range(nf)] for i in range(nf)]
    Vg = [[[np.array(v, np.float64) for v in V[i][j]] if act[i][j] else None for j in range(nf)] for i in range(nf)]
    cat = lambda lists, p: np.concatenate([l[p] for l in lists]) if lists else np.zeros(0, np.int64)
    b, brows = [None] * nf, [None] * nf
    if strategy == "fully":
        rows = [[setup_prange_without_ghosts(d) for d in test_dofs[i]] for i in range(nf)]
        for i in range(nf):
            if b_fe is not None and b_fe[i] is not None:
                b[i] = [rhs_callback(b_fe[i][p], test_dofs[i][p], rows[i][p]) for p in range(P)]
        cols = [[setup_prange_with_ghosts(trial_dofs[j][p], cat([Jg[i][j] for i in range(nf) if act[i][j]], p)) for p in range(P)]
                for j in range(nf)]
    else:
        rows = [[setup_prange_with_ghosts(test_dofs[i][p], cat([Ig[i][j] for j in range(nf) if act[i][j]], p)) for p in range(P)]
                for i in range(nf)]
        Jo = [[[trial_dofs[j][p]["l2o"][g2l(trial_dofs[j][p], Jg[i][j][p]) - 1] for p in range(P)] if act[i][j] else None
               for j in range(nf)] for i in range(nf)]
        for i in range(nf):
            for j in range(nf):
                if act[i][j]:
                    Ig[i][j], Jg[i][j], Jo[i][j], Vg[i][j] = assemble_coo_with_column_owner(Ig[i][j], Jg[i][j], Vg[i][j], rows[i], Jo[i][j])
        for i in range(nf):
            if b_fe is not None and b_fe[i] is not None:
                brows[i] = [prange_from_touched(test_dofs[i][p], touched[i][p]) for p in range(P)]
                b[i] = [rhs_callback(b_fe[i][p], test_dofs[i][p], brows[i][p]) for p in range(P)]
        cols = [[setup_prange_with_ghosts(trial_dofs[j][p], cat([Jg[i][j] for i in range(nf) if act[i][j]], p),
                                          cat([Jo[i][j] for i in range(nf) if act[i][j]], p)) for p in range(P)] for j in range(nf)]
        for i in range(nf):
            if b[i] is not None:
                b[i] = assemble_pvector(b[i], brows[i])
    out = [[None] * nf for _ in range(nf)]
    for i in range(nf):
        jb = min([j for j in range(nf) if act[i][j]], default=-1)  # the vector of block-row i is reported with its first block
        for j in range(nf):
            if not act[i][j]:
                continue
            res = []
            for p in range(P):
                li = g2l(rows[i][p], Ig[i][j][p])
                lj = g2l(cols[j][p], Jg[i][j][p])
                csr = coo_to_csr(li, lj, Vg[i][j][p], len(rows[i][p]["l2g"]), len(cols[j][p]["l2g"]))
                res.append(dict(rows=rows[i][p], cols=cols[j][p], csr=csr, b=None if (b[i] is None or j != jb) else b[i][p],
                                brows=None if (brows[i] is None or j != jb) else brows[i][p]))
            out[i][j] = res
    return out

# ------------------------------------------------------------------------------------------------
# A14  mul!  ([ext] PartitionedArrays mul!) and consistent!
# ------------------------------------------------------------------------------------------------


def consistent(x_all, cols_all):
    """owner -> ghost overwrite of a vector laid out on ``cols``."""
    owner_val = {}
    for x, c in zip(x_all, cols_all):
        nown = len(c["own_to_local"])
        owner_val.update(zip(c["l2g"][:nown].tolist(), x[:nown].tolist()))
    for x, c in zip(x_all, cols_all):
        nown = len(c["own_to_local"])
        x[nown:] = [owner_val[g] for g in c["l2g"][nown:].tolist()]
    return x_all


def mul(parts, x_all, alpha=1.0, beta=0.0, y_all=None):
    """y_own <- beta*y_own + alpha*(A_oo x_own + A_og x_ghost); only owned rows are produced."""
    x_all = consistent([x.copy() for x in x_all], [p["cols"] for p in parts])
    out = []
    for k, (p, x) in enumerate(zip(parts, x_all)):
        rowptr, colind, val = p["csr"]
        nown = len(p["rows"]["own_to_local"])
        y = np.zeros(nown)
        prod = val * x[colind]
        rid = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
        keep = rid < nown
        np.add.at(y, rid[keep], prod[keep])
        y0 = np.zeros(nown) if (y_all is None or beta == 0.0) else beta * y_all[k][:nown]
        out.append(y0 + alpha * y)
    return out
