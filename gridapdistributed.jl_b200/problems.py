"""Synthetic problems of BASELINE.json's configurations built with the host mirror (inputs only: models, spaces,
triangulations).  Used by bench.py, __graft_entry__.smoke() and the tests; nothing here computes the assembly."""
from __future__ import annotations

import numpy as np

from .fespaces import ReferenceFE, TestFESpace, TrialFESpace
from .geometry import CartesianDiscreteModel, FullyAssembledRows, Triangulation
from .parrays import DebugBackend


class Problem:
    pass


def build_problem(parts, cells, order=2, tags="boundary", ufun=None, strategy="sub", domain=None, ncomp=1, backend=None, isperiodic=None):
    """Single-field problem: CartesianDiscreteModel(parts, domain, cells), Lagrangian space of `order` with `ncomp`
    components, Dirichlet data `ufun` on `tags` (reference test/PoissonTests.jl:14-36)."""
    pr = Problem()
    D = len(parts)
    pr.backend = backend or DebugBackend(int(np.prod(parts)))
    pr.domain = domain if domain is not None else sum(([0.0, 1.0] for _ in cells), [])
    pr.model = CartesianDiscreteModel(pr.backend, parts, pr.domain, cells, isperiodic=isperiodic)
    pr.reffe = ReferenceFE("lagrangian", float, order, ncomp=ncomp)
    pr.V = TestFESpace(pr.model, pr.reffe, dirichlet_tags=tags)
    pr.U = TrialFESpace(ufun, pr.V)
    pr.strategy = strategy
    pr.trian = Triangulation(FullyAssembledRows(), pr.model) if strategy == "fully" else Triangulation(pr.model)
    pr.D, pr.order, pr.ncomp = D, order, ncomp
    return pr


def build_stokes_problem(parts, cells, strategy="sub", backend=None, ufun=None):
    """Taylor-Hood Q2^D / Q1 velocity-pressure pair, velocity Dirichlet on the whole boundary (BASELINE config 5;
    reference test/StokesOpenBoundaryTests.jl:33-49 with BlockMultiFieldStyle)."""
    D = len(cells)
    pr = Problem()
    pr.backend = backend or DebugBackend(int(np.prod(parts)))
    pr.model = CartesianDiscreteModel(pr.backend, parts, sum(([0.0, 1.0] for _ in cells), []), cells)
    pr.V = TestFESpace(pr.model, ReferenceFE("lagrangian", float, 2, ncomp=D), dirichlet_tags="boundary")
    pr.Q = TestFESpace(pr.model, ReferenceFE("lagrangian", float, 1), dirichlet_tags=None)
    if ufun is None:
        ufun = lambda x: np.stack([x[(d + 1) % D] * (1.0 - x[d]) + 0.25 * d for d in range(D)])
    pr.U = TrialFESpace(ufun, pr.V)
    pr.P = TrialFESpace(None, pr.Q)
    pr.strategy, pr.D = strategy, D
    pr.trian = Triangulation(FullyAssembledRows(), pr.model) if strategy == "fully" else Triangulation(pr.model)
    return pr


def interior_vertex_mask(m, global_interior=True):
    """1 for the vertices of local model `m` that may be moved: strictly inside the GLOBAL domain (the domain stays the unit
    box) or, with global_interior=False, strictly inside the local box (neighbouring parts see the same coordinates)."""
    ijk = m.vertex_multi_index()
    if global_interior:
        gi = ijk + np.asarray(m.cmin)[None, :]
        return np.all((gi > 0) & (gi < np.asarray(m.ncells_global)[None, :]), axis=1).astype(np.float64)
    n = np.asarray(m.ncells_local)
    return np.all((ijk > 0) & (ijk < n[None, :]), axis=1).astype(np.float64)


def vertex_perturbation(amplitude=0.1, seed=0):
    """perturb(m, xyz) for SparseMatrixAssembler(..., geometry="hex", perturb=...): every interior vertex of the GLOBAL mesh is
    moved by a pseudo-random vector of at most `amplitude` cell sizes that depends on its global index only (so all parts
    agree): general trilinear hexes, SURVEY 8d."""

    def perturb(m, xyz):
        ijk = m.vertex_multi_index() + np.asarray(m.cmin)[None, :]
        D = ijk.shape[1]
        # counter-based hash of (global vertex index, component, seed) -> uniform in [-1, 1)
        key = np.zeros(len(ijk), dtype=np.uint64)
        for d in range(D):
            key = key * np.uint64(1000003) + ijk[:, d].astype(np.uint64)
        out = np.empty_like(xyz)
        err = np.seterr(over="ignore")   # the hash wraps around on purpose
        for d in range(D):
            z = key * np.uint64(0x9E3779B97F4A7C15) + np.uint64((seed * 7919 + d + 1) * 0xBF58476D1CE4E5B9 % (1 << 64))
            z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
            z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
            out[:, d] = (z >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0
        np.seterr(**err)
        return xyz + amplitude * out * np.asarray(m.h)[None, :] * interior_vertex_mask(m)[:, None]

    return perturb
