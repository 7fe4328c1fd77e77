"""Distributed Lagrangian FE spaces on Cartesian models: INPUT PRODUCER of the assembly path.

What the hot path consumes from a space (reference FESpaces.jl:458-459, 691-743, 848-849) is
``cell_dof_ids`` per part (ncells x nd Int32, free > 0, Dirichlet < 0), the Dirichlet values, and the
dof ``PRange`` (local->global, local->owner) produced by ``generate_gids`` (FESpaces.jl:139-261).
This module restates exactly that and nothing else of the reference's space machinery.

* per-part numbering = Gridap's conforming numbering on the local box [ext, SURVEY A.6]: loop
  d = 0..D over the d-faces in face-id order (vertices: lexicographic grid nodes; d>=1 faces:
  first appearance scanning cells lexicographically and local faces in polytope order); a face whose
  Cartesian entity is in ``dirichlet_tags`` receives the next negative ids, otherwise the next
  positive ids; vector-valued spaces number the components of a node consecutively and use
  ``ldof = node + nnodes*comp`` inside the cell.
* ``generate_gids`` follows FESpaces.jl:139-261 line by line, including its three ghost-cell
  exchanges, on either backend.
"""
from __future__ import annotations

import numpy as np

from .geometry import NCube, boundary_entity_of_nodes
from .parrays import LocalIndices, PRange


class ReferenceFE:
    """``ReferenceFE(lagrangian, T, order)``; ``ncomp`` = number of components of ``T``."""

    def __init__(self, name, T=float, order=1, ncomp=None):
        assert name == "lagrangian", "only Lagrangian elements are on the accelerated path"
        assert order in (1, 2), "orders 1 and 2 are supported"
        self.order = int(order)
        if ncomp is None:
            ncomp = 1 if T in (float, np.float64) else int(T)
        self.ncomp = int(ncomp)


lagrangian = "lagrangian"


class LocalFESpace:
    """What one part knows: the serial Gridap space on its local box, reduced to id tables."""

    def __init__(self, model, reffe, dirichlet_tags):
        D = model.D
        poly = model.poly
        k, nc = reffe.order, reffe.ncomp
        self.model, self.order, self.ncomp, self.D = model, k, nc, D
        n = model.ncells_local
        ci = model.cell_multi_index()  # (ncells, D)
        ncells = len(ci)
        dims = list(range(D + 1)) if k == 2 else [0]
        self.ref_nodes = poly.q2_ref_nodes() if k == 2 else poly.q1_ref_nodes()
        nnodes_cell = len(self.ref_nodes)
        self.nd = nnodes_cell * nc
        # half-cell node grid of the local box
        hn = 2 * n + 1
        hstride = np.concatenate([[1], np.cumprod(hn[:-1])]).astype(np.int64)
        node_of_key = {}
        cell_node_keys = np.zeros((ncells, nnodes_cell), dtype=np.int64)
        col = 0
        # number the faces dimension by dimension --------------------------------------------
        face_keys_in_order = []
        nkeys = int(np.prod(hn))
        cell_key0 = (2 * ci) @ hstride                       # key of the low corner of every cell
        for d in dims:
            mids = poly.face_mid[d]  # (nlf, D) in half units
            if model.periodic_local.any():
                hpos = (2 * ci)[:, None, :] + mids[None, :, :]
                for dd in range(D):   # a locally periodic direction: the nodes of the far end ARE the nodes of the near end
                    if model.periodic_local[dd]:
                        hpos[:, :, dd] = hpos[:, :, dd] % (2 * n[dd])
                keys = hpos @ hstride  # (ncells, nlf)
            else:
                keys = cell_key0[:, None] + (mids @ hstride)[None, :]
            cell_node_keys[:, col : col + len(mids)] = keys
            col += len(mids)
            flat = keys.ravel()  # cell-major, local-face order
            # the distinct keys in first-appearance order (vertices: ascending key = lexicographic vertex ids), through a table over
            # the half-cell grid instead of a sort of all (cell, face) pairs: written backwards, the first occurrence wins
            first = np.full(nkeys, -1, dtype=np.int64)
            first[flat[::-1]] = np.arange(len(flat) - 1, -1, -1, dtype=np.int64)
            uniq = np.flatnonzero(first >= 0)
            if d > 0:
                uniq = uniq[np.argsort(first[uniq], kind="stable")]
            face_keys_in_order.append(uniq)
        all_keys = np.concatenate(face_keys_in_order)  # node id (all dims) -> half-grid key
        nnodes = len(all_keys)
        # Dirichlet tagging from the GLOBAL position of each node
        hidx = np.stack([(all_keys // hstride[d]) % hn[d] for d in range(D)], axis=1)
        hidx_glob = hidx + 2 * model.cmin[None, :]
        entity = boundary_entity_of_nodes(poly, hidx_glob, model.ncells_global, model.no_boundary)
        tags = np.zeros(poly.num_entities + 1, dtype=bool)
        for t in _resolve_tags(dirichlet_tags, poly):
            tags[t] = True
        node_is_dir = tags[entity]
        # free / Dirichlet numbering, components of a node consecutive
        node_first_free = np.zeros(nnodes, dtype=np.int64)
        node_first_dir = np.zeros(nnodes, dtype=np.int64)
        nfree_nodes = int((~node_is_dir).sum())
        ndir_nodes = int(node_is_dir.sum())
        node_first_free[~node_is_dir] = np.arange(nfree_nodes) * nc + 1
        node_first_dir[node_is_dir] = np.arange(ndir_nodes) * nc + 1
        self.num_free_dofs = nfree_nodes * nc
        self.num_dirichlet_dofs = ndir_nodes * nc
        # node lookup by key (a table over the half-cell grid)
        node_of_grid = np.full(nkeys, -1, dtype=np.int64)
        node_of_grid[all_keys] = np.arange(nnodes, dtype=np.int64)
        cell_nodes = node_of_grid[cell_node_keys]  # (ncells, nnodes_cell)
        ids = np.zeros((ncells, self.nd), dtype=np.int32)
        # per node: the id of its component 0 (free > 0, Dirichlet < 0) and the direction in which the components count
        node_id0 = np.where(node_is_dir, -node_first_dir, node_first_free).astype(np.int32)
        node_sgn = np.where(node_is_dir, -1, 1).astype(np.int32)
        cell_id0 = node_id0[cell_nodes]
        cell_sgn = node_sgn[cell_nodes] if nc > 1 else None
        for c in range(nc):
            ids[:, c * nnodes_cell : (c + 1) * nnodes_cell] = cell_id0 if c == 0 else cell_id0 + c * cell_sgn
        self.cell_dof_ids = np.ascontiguousarray(ids)
        # node coordinates of free and Dirichlet dofs (for interpolation of boundary data / sources)
        coords = model.local_origin()[None, :] + hidx * (model.h[None, :] / 2.0)
        self.free_dof_coords = np.repeat(coords[~node_is_dir], nc, axis=0)
        self.dirichlet_dof_coords = np.repeat(coords[node_is_dir], nc, axis=0)
        self.free_dof_comp = np.tile(np.arange(nc), nfree_nodes)
        self.dirichlet_dof_comp = np.tile(np.arange(nc), ndir_nodes)
        _ = node_of_key


def _resolve_tags(tags, poly):
    if tags is None:
        return []
    if isinstance(tags, str):
        tags = [tags]
    out = []
    for t in tags:
        if t == "boundary":
            out.extend(range(1, poly.num_entities))  # all but the interior
        else:
            out.append(int(t))
    return out


# ----------------------------------------------------------------------------------------------
# generate_gids (reference FESpaces.jl:139-261)
# ----------------------------------------------------------------------------------------------


class _GhostCellFetch:
    """``fetch_vector_ghost_values_cache`` + ``fetch_vector_ghost_values!`` (FESpaces.jl:130-137):
    the owner of a cell sends cell-wise data to every part that holds the cell as a ghost."""

    def __init__(self, cell_range: PRange):
        b = cell_range.backend
        self.backend = b
        req = []
        self.ghost_by_owner = []
        for ids in cell_range.indices:
            d, gb = {}, {}
            g2l = ids.ghost_to_local
            g2o = ids.ghost_to_owner
            for o in np.unique(g2o):
                sel = g2l[g2o == o]
                gb[int(o)] = sel
                d[int(o)] = ids.l2g[sel - 1]
            req.append(d)
            self.ghost_by_owner.append(gb)
        got = b.exchange(req)
        self.own_to_send = []
        for ids, r in zip(cell_range.indices, got):
            self.own_to_send.append({q: ids.global_to_local(g) for q, g in r.items()})

    def fetch(self, cell_data):
        """cell_data[k]: (ncells, nd) array; ghost-cell rows are overwritten with the owner's rows."""
        snd = [{q: cd[l - 1].ravel() for q, l in m.items()} for cd, m in zip(cell_data, self.own_to_send)]
        rcv = self.backend.exchange(snd)
        for cd, gb, r in zip(cell_data, self.ghost_by_owner, rcv):
            for o, lids in gb.items():
                cd[lids - 1] = r[o].reshape(len(lids), cd.shape[1])


def generate_gids(cell_range: PRange, cell_to_ldofs, nldofs):
    """Line-by-line restatement of reference FESpaces.jl:139-261."""
    b = cell_range.backend
    ldof_to_owner_all, nodofs = [], []
    for ids, c2l, nl in zip(cell_range.indices, cell_to_ldofs, nldofs):  # :145-163
        l2o = np.zeros(nl, dtype=np.int32)
        owner = np.broadcast_to(ids.l2o[:, None], c2l.shape)
        m = c2l > 0
        np.maximum.at(l2o, c2l[m] - 1, owner[m])
        ldof_to_owner_all.append(l2o)
        nodofs.append(int((l2o == ids.part).sum()))

    def dof_wise_to_cell_wise_fill(cwv, dwv, c2l, ids):  # :70-90 (own cells only)
        own = ids.own_to_local - 1
        sub = c2l[own]
        vals = cwv[own]
        m = sub > 0
        vals[m] = dwv[sub[m] - 1]
        cwv[own] = vals

    def cell_wise_to_dof_wise(dwv, cwv, c2l, ids):  # :92-109 (ghost cells, in order)
        gh = ids.ghost_to_local - 1
        sub = c2l[gh]
        m = sub > 0
        dwv[sub[m] - 1] = cwv[gh][m]  # repeated indices: the last assignment wins, as in the loop

    cell_parts = []  # :165-167
    for ids, c2l, l2o in zip(cell_range.indices, cell_to_ldofs, ldof_to_owner_all):
        cwv = np.full(c2l.shape, -1, dtype=np.int64)
        dof_wise_to_cell_wise_fill(cwv, l2o.astype(np.int64), c2l, ids)
        cell_parts.append(cwv)
    cache = _GhostCellFetch(cell_range)  # :177
    cache.fetch(cell_parts)  # :178
    for ids, c2l, l2o, cwv in zip(cell_range.indices, cell_to_ldofs, ldof_to_owner_all, cell_parts):
        tmp = l2o.astype(np.int64)
        cell_wise_to_dof_wise(tmp, cwv, c2l, ids)  # :180-183
        l2o[:] = tmp
    first_gdof = b.scan_exclusive(nodofs, init=1)  # :186
    ldof_to_gdof_all = []
    for ids, l2o, first in zip(cell_range.indices, ldof_to_owner_all, first_gdof):  # :189-208
        mine = l2o == ids.part
        g = np.zeros(len(l2o), dtype=np.int64)
        g[mine] = np.arange(1, int(mine.sum()) + 1) + (first - 1)
        ldof_to_gdof_all.append(g)
    cell_gdofs = []  # :211-213
    for ids, c2l, g in zip(cell_range.indices, cell_to_ldofs, ldof_to_gdof_all):
        cwv = np.full(c2l.shape, -1, dtype=np.int64)
        dof_wise_to_cell_wise_fill(cwv, g, c2l, ids)
        cell_gdofs.append(cwv)
    cache.fetch(cell_gdofs)  # :216
    for ids, c2l, cwv, g, l2o in zip(cell_range.indices, cell_to_ldofs, cell_gdofs, ldof_to_gdof_all, ldof_to_owner_all):
        gh = ids.ghost_to_local - 1  # :219-233
        sub = c2l[gh]
        cell_owner = np.broadcast_to(ids.l2o[gh][:, None], sub.shape)
        m = sub > 0
        m[m] &= l2o[sub[m] - 1] == cell_owner[m]
        g[sub[m] - 1] = cwv[gh][m]
    for ids, c2l, cwv, g in zip(cell_range.indices, cell_to_ldofs, cell_gdofs, ldof_to_gdof_all):
        dof_wise_to_cell_wise_fill(cwv, g, c2l, ids)  # :235-238
    cache.fetch(cell_gdofs)  # :240
    for ids, c2l, cwv, g in zip(cell_range.indices, cell_to_ldofs, cell_gdofs, ldof_to_gdof_all):
        cell_wise_to_dof_wise(g, cwv, c2l, ids)  # :242-245
    ngdofs = b.reduction(nodofs)  # :248
    indices = [
        LocalIndices(ng, ids.part, g, l2o)
        for ng, ids, g, l2o in zip(ngdofs, cell_range.indices, ldof_to_gdof_all, ldof_to_owner_all)
    ]
    return PRange(b, indices)  # :258


# ----------------------------------------------------------------------------------------------
# Distributed spaces
# ----------------------------------------------------------------------------------------------


class DistributedFESpace:
    """``DistributedSingleFieldFESpace`` (reference FESpaces.jl:281-300): per-part spaces + dof gids."""

    def __init__(self, model, reffe, spaces, gids, dirichlet_values=None):
        self.model, self.reffe, self.spaces, self.gids = model, reffe, spaces, gids
        self.dirichlet_values = dirichlet_values or [np.zeros(s.num_dirichlet_dofs) for s in spaces]

    def local_views(self):
        return self.spaces

    def get_free_dof_ids(self):
        return self.gids


def FESpace(model, reffe, dirichlet_tags=None, **kw):
    """``FESpace(model,reffe;dirichlet_tags)`` (reference FESpaces.jl:570-614)."""
    spaces = [LocalFESpace(m, reffe, dirichlet_tags) for m in model.models]  # :607-609
    gids = generate_gids(model.cell_gids, [s.cell_dof_ids for s in spaces], [s.num_free_dofs for s in spaces])  # :610
    return DistributedFESpace(model, reffe, spaces, gids)


TestFESpace = FESpace


def TrialFESpace(u, V: DistributedFESpace):
    """``TrialFESpace(u,V)``: Dirichlet values = ``u`` interpolated at the Dirichlet nodes."""
    vals = []
    for s in V.spaces:
        if u is None or s.num_dirichlet_dofs == 0:
            vals.append(np.zeros(s.num_dirichlet_dofs))
            continue
        r = np.asarray(u(s.dirichlet_dof_coords.T), dtype=np.float64)
        if r.ndim == 2:  # vector valued: (ncomp, n)
            r = r[s.dirichlet_dof_comp, np.arange(r.shape[1])]
        vals.append(np.ascontiguousarray(np.broadcast_to(r, (s.num_dirichlet_dofs,)), dtype=np.float64))
    return DistributedFESpace(V.model, V.reffe, V.spaces, V.gids, vals)


__all__ = ["ReferenceFE", "lagrangian", "FESpace", "TestFESpace", "TrialFESpace", "generate_gids", "DistributedFESpace"]
_ = NCube
