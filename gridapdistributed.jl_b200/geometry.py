"""Distributed Cartesian models: the INPUT PRODUCER side of the assembly path.

Mirrors (names and argument meaning) ``CartesianDiscreteModel(ranks,parts,domain,cells)`` of
reference Geometry.jl:378-411 with the cell partition of PArraysExtras.jl:5-84:

* part ``p`` (1-based, x-fastest in the part grid, PArraysExtras.jl:23,50) owns the box
  ``_local_range(p_d,np_d,n_d)`` per direction (remainder to the last parts, :72-84);
* its LOCAL box is the own box grown by one ghost layer, clipped at the domain (:24, :82-83);
* local cell ids are lexicographic (x fastest) over the local box -- the serial Gridap model of the
  local box (Geometry.jl:403-407) -- and ``local_to_owner`` follows :26-36,58-60.

The polytope tables below (vertex/edge/face order of QUAD/HEX, entity tags of Cartesian face
labelings) belong to Gridap.jl [ext], which is not under /root/reference; they are kept in ONE
place (``NCube``) so they can be corrected if a Julia depot ever shows a different order.  The C ABI
never hard-codes them: it receives ``cell_dof_ids`` and ``ref_node_coords`` as data.
"""
from __future__ import annotations

import itertools

import numpy as np

from .parrays import LocalIndices, PRange, uniform_local_range


class NCube:
    """n-cube polytope tables (Gridap QUAD / HEX) [ext, SURVEY appendix A.5].

    Vertices are lexicographic with x fastest.  d-faces are listed as tuples of vertex ids
    (0-based here).  ``faces[d]`` follows Gridap's extrusion order:
      QUAD edges  (1,2) (3,4) (1,3) (2,4)
      HEX  edges  (1,2) (3,4) (1,3) (2,4) (5,6) (7,8) (5,7) (6,8) (1,5) (2,6) (3,7) (4,8)
      HEX  faces  z=0, z=1, y=0, y=1, x=0, x=1
    """

    _EDGES = {
        1: [(0, 1)],
        2: [(0, 1), (2, 3), (0, 2), (1, 3)],
        3: [(0, 1), (2, 3), (0, 2), (1, 3), (4, 5), (6, 7), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7)],
    }
    _FACES3 = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 4, 5), (2, 3, 6, 7), (0, 2, 4, 6), (1, 3, 5, 7)]

    def __init__(self, D):
        self.D = D
        nv = 2**D
        self.vertex_coords = np.array([[(v >> d) & 1 for d in range(D)] for v in range(nv)], dtype=np.int64)
        faces = {0: [(v,) for v in range(nv)]}
        if D >= 1:
            faces[1] = list(self._EDGES[D])
        if D == 3:
            faces[2] = list(self._FACES3)
        faces[D] = [tuple(range(nv))]
        self.faces = faces
        # every face is identified by its midpoint in half-cell units (0,1,2 per direction)
        self.face_mid = {
            d: np.array([self.vertex_coords[list(f)].sum(0) * 2 // len(f) for f in fs], dtype=np.int64)
            for d, fs in faces.items()
        }
        # entity id (1-based, Cartesian face labeling): vertices, then edges, ..., then interior
        self.entity_of_mid = {}
        e = 0
        for d in range(D + 1):
            for m in self.face_mid[d]:
                e += 1
                self.entity_of_mid[tuple(int(x) for x in m)] = e
        self.num_entities = e

    def q2_ref_nodes(self):
        """Reference coordinates of the order-2 Lagrangian nodes in Gridap's local dof order:
        vertices, edge interiors, face interiors, cell interior."""
        mids = [self.face_mid[d] for d in range(self.D + 1)]
        return np.concatenate(mids, 0).astype(np.float64) / 2.0

    def q1_ref_nodes(self):
        return self.vertex_coords.astype(np.float64)


class CartesianLocalModel:
    """The serial ``CartesianDiscreteModel(desc,cmin,cmax)`` of one part's local box."""

    def __init__(self, D, origin, h, ncells_global, cmin, ncells_local, cell_owner, part, periodic_local=None, no_boundary=None):
        self.D = D
        # periodic_local[d]: the LOCAL model is periodic in direction d (periodic and not partitioned: the nodes of its two ends
        # are identified); no_boundary[d]: direction d has no boundary at all (periodic, partitioned or not) -- reference
        # Geometry.jl:413-459 (`remove_boundary`, local periodicity only in the directions that are not partitioned)
        self.periodic_local = np.zeros(D, dtype=bool) if periodic_local is None else np.asarray(periodic_local, dtype=bool)
        self.no_boundary = np.zeros(D, dtype=bool) if no_boundary is None else np.asarray(no_boundary, dtype=bool)
        self.origin = np.asarray(origin, dtype=np.float64)  # global origin
        self.h = np.asarray(h, dtype=np.float64)
        self.ncells_global = np.asarray(ncells_global, dtype=np.int64)
        self.cmin = np.asarray(cmin, dtype=np.int64)  # 0-based global index of the first local cell
        self.ncells_local = np.asarray(ncells_local, dtype=np.int64)
        self.cell_owner = cell_owner  # int32 per local cell (1-based parts)
        self.part = part
        self.poly = NCube(D)

    @property
    def num_cells(self):
        return int(np.prod(self.ncells_local))

    def local_origin(self):
        return self.origin + self.cmin * self.h

    def cell_multi_index(self):
        """(ncells, D) local multi-index of each local cell, lexicographic x fastest."""
        grids = np.meshgrid(*[np.arange(n) for n in self.ncells_local[::-1]], indexing="ij")
        return np.stack([g.ravel() for g in grids[::-1]], axis=1)

    def vertex_multi_index(self):
        """(nvertices, D) local multi-index of each local vertex, lexicographic x fastest."""
        nv = self.ncells_local + 1
        grids = np.meshgrid(*[np.arange(n) for n in nv[::-1]], indexing="ij")
        return np.stack([g.ravel() for g in grids[::-1]], axis=1)

    def vertex_coordinates(self):
        return self.local_origin() + self.vertex_multi_index() * self.h

    def cell_vertex_ids(self):
        """(ncells, 2^D) 1-based local vertex ids, x fastest within the cell."""
        nv = self.ncells_local + 1
        ci = self.cell_multi_index()
        stride = np.concatenate([[1], np.cumprod(nv[:-1])])
        base = (ci * stride).sum(1)
        offs = (self.poly.vertex_coords * stride).sum(1)
        return (base[:, None] + offs[None, :] + 1).astype(np.int32)


class DistributedCartesianDiscreteModel:
    """``CartesianDiscreteModel(ranks, parts, domain, cells)`` (reference Geometry.jl:378-411)."""

    def __init__(self, ranks, parts, domain, cells, isperiodic=None):
        self.backend = ranks
        self.isperiodic = tuple(bool(p) for p in isperiodic) if isperiodic is not None else tuple(False for _ in parts)
        self.parts = tuple(int(p) for p in parts)
        self.D = D = len(self.parts)
        assert len(cells) == D and len(domain) == 2 * D
        assert int(np.prod(self.parts)) == ranks.nparts
        self.cells = tuple(int(c) for c in cells)
        dom = np.asarray(domain, dtype=np.float64).reshape(D, 2)
        self.origin = dom[:, 0].copy()
        self.h = (dom[:, 1] - dom[:, 0]) / np.asarray(self.cells)
        self.models = []
        cell_indices = []
        nglob = int(np.prod(self.cells))
        for part in ranks.parts_here:
            m, ids = self._build_part(part, nglob)
            self.models.append(m)
            cell_indices.append(ids)
        self.cell_gids = PRange(ranks, cell_indices)

    def _build_part(self, part, nglob):
        D = self.D
        # CartesianIndices(np)[rank]: x fastest (PArraysExtras.jl:23)
        p = []
        r = part - 1
        for d in range(D):
            p.append(r % self.parts[d] + 1)
            r //= self.parts[d]
        # a periodic AND partitioned direction gets a ghost cell on either side that wraps around (Geometry.jl:436-449:
        # local_range(...,ghost,global_isperiodic)); a periodic direction with one part is handled by the local model
        wrap = [self.isperiodic[d] and self.parts[d] != 1 for d in range(D)]
        loc = [uniform_local_range(p[d], self.parts[d], self.cells[d], ghost=True, periodic=wrap[d]) for d in range(D)]
        # per direction owner of every local index (PArraysExtras.jl:26-36)
        owners_d = []
        for d in range(D):
            bounds = [uniform_local_range(q, self.parts[d], self.cells[d]) for q in range(1, self.parts[d] + 1)]
            idx = (np.arange(loc[d][0], loc[d][1] + 1) - 1) % self.cells[d] + 1   # wrapped global index of every local index
            o = np.zeros(len(idx), dtype=np.int64)
            for q, (a, b) in enumerate(bounds):
                o[(idx >= a) & (idx <= b)] = q + 1
            owners_d.append(o)
        nloc = np.array([loc[d][1] - loc[d][0] + 1 for d in range(D)], dtype=np.int64)
        cmin = np.array([loc[d][0] - 1 for d in range(D)], dtype=np.int64)
        grids = np.meshgrid(*[np.arange(n) for n in nloc[::-1]], indexing="ij")
        ci = np.stack([g.ravel() for g in grids[::-1]], axis=1)  # lexicographic, x fastest
        # owner part id = LinearIndices(np)[owners...] (PArraysExtras.jl:58-60)
        owner = np.zeros(len(ci), dtype=np.int64)
        gid = np.zeros(len(ci), dtype=np.int64)
        pstride, gstride = 1, 1
        for d in range(D):
            owner += (owners_d[d][ci[:, d]] - 1) * pstride
            gid += ((ci[:, d] + cmin[d]) % self.cells[d]) * gstride
            pstride *= self.parts[d]
            gstride *= self.cells[d]
        owner = (owner + 1).astype(np.int32)
        gid = gid + 1
        model = CartesianLocalModel(D, self.origin, self.h, self.cells, cmin, nloc, owner, part,
                                    periodic_local=[self.isperiodic[d] and self.parts[d] == 1 for d in range(D)],
                                    no_boundary=list(self.isperiodic))
        ids = LocalIndices(nglob, part, gid, owner)
        return model, ids

    def local_views(self):
        return self.models


def CartesianDiscreteModel(ranks, parts, domain, cells, isperiodic=None):
    """``CartesianDiscreteModel(ranks,parts,domain,cells;isperiodic=...)`` (reference Geometry.jl:378-459)."""
    return DistributedCartesianDiscreteModel(ranks, parts, domain, cells, isperiodic)


# ----------------------------------------------------------------------------------------------
# Triangulations: which cells each part integrates (reference Geometry.jl:797-805)
# ----------------------------------------------------------------------------------------------


class SubAssembledRows:
    """Integrate owned cells only; ghost rows migrate to their owner (reference Algebra.jl:398)."""

    code = 0


class FullyAssembledRows:
    """Integrate owned+ghost cells; keep owned rows only; no matrix communication (Algebra.jl:397)."""

    code = 1


class DistributedTriangulation:
    def __init__(self, model, cell_lids):
        self.model = model
        self.cell_lids = cell_lids  # per part: int32 1-based local cell ids, ascending


def Triangulation(*args):
    """``Triangulation(model)`` removes ghost cells; ``Triangulation(FullyAssembledRows(),model)`` /
    ``with_ghost`` keeps them (reference Geometry.jl:797-805, 841-847)."""
    if len(args) == 1:
        portion, model = SubAssembledRows(), args[0]
    else:
        portion, model = args
    lids = []
    for m, ids in zip(model.models, model.cell_gids.indices):
        if isinstance(portion, FullyAssembledRows) or portion == "with_ghost":
            lids.append(np.arange(1, m.num_cells + 1, dtype=np.int32))
        else:
            lids.append(ids.own_to_local.copy())
    return DistributedTriangulation(model, lids)


class DistributedBoundaryTriangulation:
    """``Boundary(model,tags=...)`` (reference Geometry.jl:684-767): per part the boundary facets of the LOCAL cells whose
    Cartesian entity carries one of the tags, as (1-based local cell id, facet of the n-cube in Gridap's order: HEX z=0, z=1,
    y=0, y=1, x=0, x=1; QUAD y=0, y=1, x=0, x=1).  Which of them are integrated follows from the body triangulation of the
    form (owned cells for SubAssembledRows, owned + ghost for FullyAssembledRows)."""

    def __init__(self, model, tags):
        self.model = model
        self.cell_lids, self.lfaces = [], []
        for m in model.models:
            D = m.D
            want = np.zeros(m.poly.num_entities + 1, dtype=bool)
            for t in ([tags] if isinstance(tags, (int, str)) else list(tags)):
                if t == "boundary":
                    want[1:m.poly.num_entities] = True
                else:
                    want[int(t)] = True
            ci = m.cell_multi_index() + m.cmin[None, :]
            cells, lfaces = [], []
            for lf in range(2 * D):
                axis, side = D - 1 - lf // 2, lf % 2
                if m.no_boundary[axis]:
                    continue
                mid = [1] * D
                mid[axis] = 2 * side
                if not want[m.poly.entity_of_mid[tuple(mid)]]:
                    continue
                sel = np.flatnonzero(ci[:, axis] == (m.ncells_global[axis] - 1 if side else 0))
                cells.append(sel + 1)
                lfaces.append(np.full(len(sel), lf))
            cells = np.concatenate(cells) if cells else np.zeros(0, dtype=np.int64)
            lfaces = np.concatenate(lfaces) if lfaces else np.zeros(0, dtype=np.int64)
            order = np.lexsort((lfaces, cells))     # by cell, then by facet
            self.cell_lids.append(cells[order].astype(np.int32))
            self.lfaces.append(lfaces[order].astype(np.int32))


def Boundary(model, tags="boundary"):
    return DistributedBoundaryTriangulation(model, tags)


def facet_points(cell_coords, lfaces, degree):
    """Physical coordinates (nf, nqf, D) and outward unit normals (nf, nqf, D) of the facet quadrature points the library
    integrates on (graft_neumann_set): tensor Gauss-Legendre of n = ceil((degree+1)/2) points over the facet's free axes,
    lowest axis fastest.  This is where the caller evaluates its boundary data g(x, n) -- the lazy CellField of the
    reference (test/PoissonTests.jl:29: ``g = n_Γn⋅∇(u)``)."""
    X = np.asarray(cell_coords, dtype=np.float64)
    nf, nv, D = X.shape
    n = (int(degree) + 2) // 2
    x1 = (np.polynomial.legendre.leggauss(n)[0] + 1.0) / 2.0
    nqf = n ** (D - 1)
    axis = D - 1 - np.asarray(lfaces) // 2
    side = np.asarray(lfaces) % 2
    xi = np.zeros((nf, nqf, D))
    for a in range(D):
        sel = np.flatnonzero(axis == a)
        if len(sel) == 0:
            continue
        xi[sel, :, a] = side[sel, None]
        for o, d in enumerate([d for d in range(D) if d != a]):
            xi[np.ix_(sel, np.arange(nqf), [d])] = x1[(np.arange(nqf) // n**o) % n][None, :, None]
    bits = np.array([[(v >> d) & 1 for d in range(D)] for v in range(nv)])
    fac = np.where(bits[None, None, :, :] == 1, xi[:, :, None, :], 1.0 - xi[:, :, None, :])       # (nf, nqf, nv, D)
    sgn = np.where(bits == 1, 1.0, -1.0)
    N = fac.prod(axis=3)
    xq = np.einsum("fqv,fvd->fqd", N, X)
    J = np.zeros((nf, nqf, D, D))
    for a in range(D):
        dN = sgn[None, None, :, a] * np.prod(np.delete(fac, a, axis=3), axis=3)
        J[:, :, :, a] = np.einsum("fqv,fvd->fqd", dN, X)
    cof = np.linalg.det(J)[..., None, None] * np.swapaxes(np.linalg.inv(J), -1, -2)                # det(J) J^{-T}
    nrm = cof[np.arange(nf)[:, None], np.arange(nqf)[None, :], :, axis[:, None]] * np.where(side == 1, 1.0, -1.0)[:, None, None]
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    return xq, nrm


class Measure:
    """``Measure(Ω,degree)``: tensor Gauss-Legendre, n = ceil((degree+1)/2) points per direction."""

    def __init__(self, trian, degree):
        self.trian = trian
        self.degree = int(degree)


def boundary_entity_of_nodes(poly: NCube, half_idx_global, ncells_global, no_boundary=None):
    """Entity id (Cartesian face-labeling tag, 1-based) of nodes given by GLOBAL half-cell indices.

    flag per direction: 0 = on the low boundary, 2 = on the high boundary, 1 = interior.  The
    entity is the face of the unit n-cube with that midpoint (2-D corners 1-4, edges
    5-8, interior 9; 3-D corners 1-8, edges 9-20, faces 21-26, interior 27).
    """
    D = poly.D
    n2 = 2 * np.asarray(ncells_global, dtype=np.int64)
    flag = np.ones(half_idx_global.shape, dtype=np.int64)
    flag[half_idx_global == 0] = 0
    flag[half_idx_global == n2[None, :]] = 2
    if no_boundary is not None:   # periodic directions have no boundary
        flag[:, np.asarray(no_boundary, dtype=bool)] = 1
    code = np.zeros(len(half_idx_global), dtype=np.int64)
    for d in range(D):
        code += flag[:, d] * 3**d
    table = np.zeros(3**D, dtype=np.int32)
    for mid, e in poly.entity_of_mid.items():
        c = sum(mid[d] * 3**d for d in range(D))
        table[c] = e
    return table[code]


__all__ = [
    "NCube",
    "CartesianDiscreteModel",
    "DistributedCartesianDiscreteModel",
    "Triangulation",
    "Measure",
    "Boundary",
    "facet_points",
    "SubAssembledRows",
    "FullyAssembledRows",
    "boundary_entity_of_nodes",
]
_ = itertools
