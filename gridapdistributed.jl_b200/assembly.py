"""Host mirror of the reference's assembler interface for the accelerated path.

Same names, argument meaning and error behaviour as the reference (Julia is absent from this image,
so the host side above the C ABI is Python; the Julia shim with the same calls is julia/GraftAssembly.jl):

  SparseMatrixAssembler(U, V[, strategy])           reference FESpaces.jl:819-862
  assemble_matrix_and_vector(form, assem)            Gridap generic driver -> FESpaces.jl:763-798, Algebra.jl:593-646
  assemble_matrix / assemble_vector                  same, one output
  allocate_matrix_and_vector / assemble_matrix_and_vector_b (the reference's `assemble_matrix_and_vector!`)
  AffineFEOperator(form, U, V, assem)                reference test/PoissonTests.jl:41
  PSparseMatrix / PVector, mul(c, A, b, alpha, beta) PartitionedArrays mul! behind BlockPartitionedArrays.jl:330-351

Weak forms come from a fixed catalogue (``Poisson``, ``Mass``, ``LinearElasticity``, ``StokesTH``) that
names what ``∫(∇(v)⋅∇(u))dΩ`` etc. mean; every flop and every byte of the assembly runs in
lib/libgraft.so on the GPU.  Nothing here computes: without the library or a device the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import libgraft as L
from .geometry import FullyAssembledRows, Measure, SubAssembledRows, facet_points
from .parrays import OwnAndGhostIndices, PRange

# ----------------------------------------------------------------------------------------------
# weak-form catalogue
# ----------------------------------------------------------------------------------------------


class GraftForm:
    form_id = 0
    params: tuple = ()

    def __init__(self, dΩ: Measure, source=None, extra_cellvec=None, neumann=None):
        self.dΩ = dΩ
        self.source = source  # None | float | sequence(ncomp) | ("nodal", free_values_per_part, dirichlet_values_per_part)
        self.extra_cellvec = extra_cellvec  # per part (ncells_integrated, nd) host-computed additive term
        # boundary term  + ∫( v*g )dΓ  of the right-hand side (reference test/PoissonTests.jl:39): (Measure(Boundary(model,tags),degree),
        # g) with g(x, n) -> values at the points x (D, npts) with outward unit normals n (D, npts); a dict {field: (dΓ, g)} for
        # block systems
        self.neumann = neumann


class Poisson(GraftForm):
    """a(u,v) = ∫ ∇(v)⋅∇(u) dΩ ,  l(v) = ∫ v f dΩ   (reference test/PoissonTests.jl:38-39)"""

    form_id = L.FORM_POISSON


class Mass(GraftForm):
    """a(u,v) = ∫ v u dΩ"""

    form_id = L.FORM_MASS


class LinearElasticity(GraftForm):
    """a(u,v) = ∫ ε(v) ⊙ σ(ε(u)) dΩ , σ(ε) = λ tr(ε) I + 2 μ ε"""

    form_id = L.FORM_ELASTICITY

    def __init__(self, dΩ, lam, mu, source=None, extra_cellvec=None, neumann=None):
        super().__init__(dΩ, source, extra_cellvec, neumann)
        self.params = (float(lam), float(mu))


class StokesTH(GraftForm):
    """a((u,p),(v,q)) = ∫ ν ∇(v)⊙∇(u) - (∇⋅v) p - q (∇⋅u) dΩ  (reference test/StokesOpenBoundaryTests.jl:44-49)"""

    form_id = L.FORM_STOKES

    def __init__(self, dΩ, nu=1.0, source=None, extra_cellvec=None, neumann=None):
        super().__init__(dΩ, source, extra_cellvec, neumann)
        self.params = (float(nu),)


class PLaplacian(GraftForm):
    """State-dependent forms of reference test/PLaplacianTests.jl:24-31 at the state ``uh`` (per part: values on the free dofs):
    matrix = jacobian j(uh,du,v) = ∫ ∇(v)⋅(dσ∘(∇(du),∇(uh))) dΩ, vector = residual r(uh,v) = ∫ ∇(v)⋅(σ∘∇(uh)) - v*f dΩ with
    σ(∇u) = (1+∇u⋅∇u)∇u -- what ``residual_and_jacobian!(b,A,op,uh)`` assembles (test/PLaplacianTests.jl:45-52)."""

    form_id = L.FORM_PLAPLACIAN

    def __init__(self, dΩ, uh, source=None):
        super().__init__(dΩ, source, None)
        self.uh = uh


class CellArrays(GraftForm):
    """Hook 1, "bring your own cell arrays": the ``data`` of ``collect_cell_matrix_and_vector(trial,test,a,l)`` (reference
    FESpaces.jl:691-743) already evaluated -- ``mats[(i,j)][part]``: (ncells, nd_i, nd_j) cell matrices of block (i,j) (a
    missing key = a block the form does not touch, the ``touched`` of Gridap's ArrayBlock), ``vecs[i][part]``: (ncells, nd_i)
    Dirichlet-lifted cell vectors of block row i.  ``assemble_matrix_and_vector(data, assem)`` is then the reference's
    ``assemble_matrix_and_vector(assem, data)`` (FESpaces.jl:763-798; blocks: MultiField.jl:473-560)."""

    form_id = L.FORM_USER

    def __init__(self, dΩ, mats, vecs=None):
        super().__init__(dΩ, None, None)
        if not isinstance(mats, dict):
            mats = {(0, 0): mats}
        if vecs is not None and not isinstance(vecs, dict):
            vecs = {0: vecs}
        self.mats, self.vecs = mats, (vecs or {})
        nf = 1 + max(max(i, j) for i, j in mats)
        self.params = tuple(1.0 if (i, j) in mats else 0.0 for i in range(nf) for j in range(nf))


def collect_cell_matrix_and_vector(dΩ, mats, vecs=None):
    return CellArrays(dΩ, mats, vecs)


# ----------------------------------------------------------------------------------------------
# communicator handling: DebugBackend -> local comm, DistBackend -> NCCL comm
# ----------------------------------------------------------------------------------------------


class _Comm:
    def __init__(self, backend, device=None):
        import torch

        lib = L.load()
        self.lib = lib
        self.backend = backend
        self.handle = L.c_vp()
        self.ctxs = []
        if backend.name == "debug":
            L.check(lib.graft_comm_create_local(backend.nparts, C.byref(self.handle)))
            ndev = max(torch.cuda.device_count(), 1)
            for p in backend.parts_here:
                ctx = L.c_vp()
                dev = device if device is not None else 0
                _ = ndev
                L.check(lib.graft_ctx_create(self.handle, p, dev, C.byref(ctx)))
                self.ctxs.append(ctx)
        else:
            import torch.distributed as dist

            uid = np.zeros(128, dtype=np.uint8)
            if backend.rank == 0:
                L.check(lib.graft_nccl_unique_id(L.ptr(uid)))
            box = [uid.tobytes()]
            dist.broadcast_object_list(box, src=0, group=backend.group)
            uid = np.frombuffer(box[0], dtype=np.uint8).copy()
            dev = device if device is not None else torch.cuda.current_device()
            L.check(lib.graft_comm_create_nccl(backend.rank, backend.nparts, L.ptr(uid), dev, C.byref(self.handle)))
            ctx = L.c_vp()
            L.check(lib.graft_ctx_create(self.handle, backend.rank + 1, dev, C.byref(ctx)))
            self.ctxs.append(ctx)

    def close(self):
        if self.handle:
            for ctx in self.ctxs:
                self.lib.graft_ctx_destroy(ctx)
            self.lib.graft_comm_destroy(self.handle)
            self.handle = None
            self.ctxs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
# distributed outputs
# ----------------------------------------------------------------------------------------------


class PVector:
    """``PVector(values, partition(rows))``: per part ``local_length(rows)`` values, own first."""

    def __init__(self, values, index_partition: PRange):
        self.vector_partition = values
        self.index_partition = index_partition

    def own_values(self):
        return [v[: ids.own_length] for v, ids in zip(self.vector_partition, self.index_partition.indices)]


class PSparseMatrix:
    """``PSparseMatrix(values, partition(rows), partition(cols))`` (reference Algebra.jl:1013-1015).

    The values live on the GPU; ``matrix_partition`` downloads them on demand as scipy CSR matrices
    (0-based) -- the analogue of ``SparseMatrixCSR{0,Float64,Int}`` local matrices.
    """

    def __init__(self, assem, bi, bj, rows: PRange, cols: PRange):
        self.assem, self.bi, self.bj = assem, bi, bj
        self.row_partition, self.col_partition = rows, cols
        self._pattern = None
        # the values are a VIEW of the assembler's device buffers: a later matrix assembly on the same assembler overwrites
        # them.  The generation stamp turns a read through a stale handle into an error instead of silently wrong values
        # (the in-place variant assemble_matrix_and_vector_b re-stamps the matrix it is given).
        self.generation = assem._matrix_generation

    def _check_fresh(self):
        if self.generation != self.assem._matrix_generation:
            raise RuntimeError("this PSparseMatrix was overwritten by a later matrix assembly on the same assembler: copy its values "
                               "(csr_arrays) before re-assembling, or re-assemble in place with assemble_matrix_and_vector_b")

    def csr_arrays(self, values_only=False):
        """Per part (rowptr, colind, nzval) in the assembler's index base."""
        self._check_fresh()
        lib, out = self.assem.comm.lib, []
        for k, ctx in enumerate(self.assem.comm.ctxs):
            m, n, nnz = L.c_i64(), L.c_i64(), L.c_i64()
            L.check(lib.graft_csr_query(ctx, self.bi, self.bj, C.byref(m), C.byref(n), C.byref(nnz)))
            val = np.empty(nnz.value, dtype=np.float64)
            if values_only and self._pattern is not None:
                L.check(lib.graft_csr_get_values(ctx, self.bi, self.bj, L.ptr(val)))
                out.append((self._pattern[k][0], self._pattern[k][1], val))
                continue
            rowptr = np.empty(m.value + 1, dtype=np.int64)
            colind = np.empty(nnz.value, dtype=np.int64)
            L.check(lib.graft_csr_get(ctx, self.bi, self.bj, L.ptr(rowptr), L.ptr(colind), L.ptr(val)))
            out.append((rowptr, colind, val))
        self._pattern = [(o[0], o[1]) for o in out]
        return out

    @property
    def matrix_partition(self):
        import scipy.sparse as sp

        base = self.assem.index_base
        mats = []
        for (rowptr, colind, val), r, c in zip(self.csr_arrays(), self.row_partition.indices, self.col_partition.indices):
            mats.append(sp.csr_matrix((val, colind - base, rowptr - base), shape=(r.local_length, c.local_length)))
        return mats


class AffineFEOperator:
    """``AffineFEOperator(form, U, V, assem)`` (reference test/PoissonTests.jl:41): assembles A, b."""

    def __init__(self, form, U, V, assem=None):
        self.assem = assem if assem is not None else SparseMatrixAssembler(U, V)
        self.trial, self.test = U, V
        self.matrix, self.vector = assemble_matrix_and_vector(form, self.assem)

    def get_matrix(self):
        return self.matrix

    def get_vector(self):
        return self.vector


# ----------------------------------------------------------------------------------------------
# the assembler
# ----------------------------------------------------------------------------------------------


class GraftSparseMatrixAssembler:
    """Drop-in for ``DistributedSparseMatrixAssembler`` (reference FESpaces.jl:746-753)."""

    def __init__(self, trial, test, strategy=None, index_base=0, geometry="cartesian", device=None, perturb=None):
        trials = trial if isinstance(trial, (list, tuple)) else [trial]
        tests = test if isinstance(test, (list, tuple)) else [test]
        assert len(trials) == len(tests)
        self.trials, self.tests = list(trials), list(tests)
        self.strategy = strategy if strategy is not None else SubAssembledRows()  # default of FESpaces.jl:824
        self.index_base = int(index_base)
        self.model = trials[0].model
        self.backend = self.model.backend
        self.comm = _Comm(self.backend, device)
        self.geometry = geometry
        self._symbolic_key = None
        self._cells_key = None
        self._matrix_generation = 0
        self._xyz = [None] * len(self.model.models)   # node coordinates handed to graft_mesh_set_hex (None: the Cartesian ones)
        self.rows = self.cols = None
        lib = self.comm.lib
        for k, ctx in enumerate(self.comm.ctxs):
            m = self.model.models[k]
            if geometry == "cartesian":
                L.check(lib.graft_mesh_set_cartesian(ctx, m.D, L.ptr(np.ascontiguousarray(m.local_origin())),
                                                     L.ptr(np.ascontiguousarray(m.h)), L.ptr(np.ascontiguousarray(m.ncells_local))))
            else:
                xyz = np.ascontiguousarray(m.vertex_coordinates(), dtype=np.float64)
                if perturb is not None:
                    xyz = np.ascontiguousarray(perturb(m, xyz))
                    self._xyz[k] = xyz
                cn = np.ascontiguousarray(m.cell_vertex_ids(), dtype=np.int32)
                L.check(lib.graft_mesh_set_hex(ctx, m.D, len(xyz), L.ptr(xyz), len(cn), L.ptr(cn)))
            for f, (U, V) in enumerate(zip(self.trials, self.tests)):
                assert U.spaces is V.spaces or U.spaces == V.spaces, "trial and test spaces must share the FE space"
                s = U.spaces[k]
                ids = U.gids.indices[k]
                dv = np.ascontiguousarray(U.dirichlet_values[k], dtype=np.float64)
                L.check(lib.graft_space_set(ctx, f, s.order, s.ncomp, s.cell_dof_ids.shape[0], L.ptr(s.cell_dof_ids), s.num_free_dofs,
                                            L.ptr(ids.l2g), L.ptr(ids.l2o), ids.n_global, s.num_dirichlet_dofs,
                                            L.ptr(dv) if len(dv) else None, L.ptr(np.ascontiguousarray(s.ref_nodes))))

    # getters of FESpaces.jl:757-761
    def get_rows(self):
        return self.rows

    def get_cols(self):
        return self.cols

    def get_assembly_strategy(self):
        return self.strategy

    def close(self):
        self.comm.close()

    # -- internals ---------------------------------------------------------------------------
    def _set_form(self, form: GraftForm):
        lib = self.comm.lib
        trian = form.dΩ.trian
        params = np.asarray(form.params, dtype=np.float64)
        for k, ctx in enumerate(self.comm.ctxs):
            lids = np.ascontiguousarray(trian.cell_lids[k], dtype=np.int32)
            key = (id(trian), len(lids))
            if self._cells_key is None or self._cells_key[k] != key:
                L.check(lib.graft_cells_set(ctx, len(lids), L.ptr(lids)))
            L.check(lib.graft_form_set(ctx, form.form_id, L.ptr(params) if len(params) else None, len(params), form.dΩ.degree))
            for f, U in enumerate(self.trials):
                s = U.spaces[k]
                src = form.source[f] if isinstance(form.source, dict) else (form.source if f == 0 else None)
                if src is None:
                    L.check(lib.graft_source_set(ctx, f, L.SOURCE_NONE, None, None))
                elif isinstance(src, tuple) and src[0] == "nodal":
                    fv = np.ascontiguousarray(src[1][k], dtype=np.float64)
                    dv = np.ascontiguousarray(src[2][k], dtype=np.float64)
                    L.check(lib.graft_source_set(ctx, f, L.SOURCE_NODAL, L.ptr(fv), L.ptr(dv) if len(dv) else None))
                else:
                    cv = np.ascontiguousarray(np.broadcast_to(np.asarray(src, dtype=np.float64), (s.ncomp,)))
                    L.check(lib.graft_source_set(ctx, f, L.SOURCE_CONST, L.ptr(cv), None))
                if form.form_id == L.FORM_PLAPLACIAN and f == 0:
                    uh = np.ascontiguousarray(form.uh[k], dtype=np.float64)
                    assert len(uh) == s.num_free_dofs
                    L.check(lib.graft_state_set(ctx, 0, L.ptr(uh) if len(uh) else None))
                ex = None
                if form.extra_cellvec is not None:
                    ex = form.extra_cellvec[f][k] if isinstance(form.extra_cellvec, dict) else (form.extra_cellvec[k] if f == 0 else None)
                L.check(lib.graft_extra_cellvec_set(ctx, f, L.ptr(np.ascontiguousarray(ex, dtype=np.float64)) if ex is not None else None))
                nm = getattr(form, "neumann", None)
                nm = nm.get(f) if isinstance(nm, dict) else (nm if f == 0 else None)
                self._set_neumann(ctx, k, f, lids, nm)
        self._cells_key = [(id(trian), len(trian.cell_lids[k])) for k in range(len(self.comm.ctxs))]

    def _set_neumann(self, ctx, k, f, lids, nm):
        """graft_neumann_set: the facets of Boundary(model,tags) whose parent cell is integrated, and g at their quadrature points."""
        lib = self.comm.lib
        if nm is None:
            L.check(lib.graft_neumann_set(ctx, f, 0, None, None, 0, None))
            return
        dΓ, gfun = nm
        Γ, m = dΓ.trian, self.model.models[k]
        pos = np.searchsorted(lids, Γ.cell_lids[k])
        keep = (pos < len(lids)) & (lids[np.minimum(pos, len(lids) - 1)] == Γ.cell_lids[k]) if len(lids) else np.zeros(0, dtype=bool)
        cells = np.ascontiguousarray(pos[keep], dtype=np.int32)
        lfaces = np.ascontiguousarray(Γ.lfaces[k][keep], dtype=np.int32)
        if len(cells) == 0:
            L.check(lib.graft_neumann_set(ctx, f, 0, None, None, 0, None))
            return
        xyz = self._xyz[k] if self._xyz[k] is not None else m.vertex_coordinates()
        X = xyz[m.cell_vertex_ids()[lids[cells] - 1] - 1]
        xq, nrm = facet_points(X, lfaces, dΓ.degree)
        nf, nqf, D = xq.shape
        ncomp = self.trials[f].spaces[k].ncomp
        vals = np.asarray(gfun(xq.reshape(-1, D).T, nrm.reshape(-1, D).T), dtype=np.float64)
        vals = np.broadcast_to(vals, (ncomp, nf * nqf)) if vals.ndim < 2 else vals.reshape(ncomp, nf * nqf)
        g = np.ascontiguousarray(vals.T.reshape(nf, nqf, ncomp))
        L.check(lib.graft_neumann_set(ctx, f, nf, L.ptr(cells), L.ptr(lfaces), dΓ.degree, L.ptr(g)))

    def _prange(self, which, blk):
        lib, idx = self.comm.lib, []
        for ctx, p in zip(self.comm.ctxs, self.backend.parts_here):
            no, ng, nglob = L.c_i64(), L.c_i64(), L.c_i64()
            L.check(lib.graft_prange_query(ctx, which, blk, C.byref(no), C.byref(ng), C.byref(nglob)))
            o2g = np.empty(no.value, dtype=np.int64)
            g2g = np.empty(ng.value, dtype=np.int64)
            g2o = np.empty(ng.value, dtype=np.int32)
            L.check(lib.graft_prange_get(ctx, which, blk, L.ptr(o2g), L.ptr(g2g), L.ptr(g2o)))
            idx.append(OwnAndGhostIndices(nglob.value, p, o2g, g2g, g2o))
        return PRange(self.backend, idx)

    def _symbolic(self, form):
        key = (form.form_id == L.FORM_STOKES, form.params if form.form_id == L.FORM_USER else None, id(form.dΩ.trian), self.strategy.code)
        if self._symbolic_key == key:
            return
        L.check(self.comm.lib.graft_symbolic(self.comm.handle, self.strategy.code, self.index_base))
        nf = len(self.trials)
        self.rows = [self._prange(0, f) for f in range(nf)]
        self.cols = [self._prange(1, f) for f in range(nf)]
        self.brows = [self._prange(2, f) for f in range(nf)]
        self._symbolic_key = key

    def _numeric(self, what, form=None):
        if what & 1:
            self._matrix_generation += 1
        if isinstance(form, CellArrays):
            return self._scatter(form, what)
        L.check(self.comm.lib.graft_numeric(self.comm.handle, what))

    def _scatter(self, data, what):
        """graft_scatter_cellmats block by block; the vector of block row i travels with the first block of that row."""
        lib = self.comm.lib
        keep = []

        def per_part(arrs):
            a = [np.ascontiguousarray(x, dtype=np.float64) for x in arrs]
            keep.append(a)
            return L.ptr_array(a)

        done = set()
        for (i, j) in sorted(data.mats):
            mats = per_part(data.mats[(i, j)]) if what & 1 else None
            vecs = None
            if (what & 2) and i in data.vecs and i not in done:
                vecs = per_part(data.vecs[i])
                done.add(i)
            if mats is not None or vecs is not None:
                L.check(lib.graft_scatter_cellmats(self.comm.handle, i, j, mats, vecs))

    def _wrap(self):
        nf = len(self.trials)
        mats = [[PSparseMatrix(self, i, j, self.rows[i], self.cols[j]) for j in range(nf)] for i in range(nf)]
        return mats[0][0] if nf == 1 else mats

    def _vectors(self):
        lib, nf = self.comm.lib, len(self.trials)
        out = []
        for f in range(nf):
            vals = []
            for ctx, ids in zip(self.comm.ctxs, self.brows[f].indices):
                b = np.empty(ids.local_length, dtype=np.float64)
                L.check(lib.graft_vec_get(ctx, f, L.ptr(b)))
                vals.append(b)
            out.append(PVector(vals, self.brows[f]))
        return out[0] if nf == 1 else out

    def timers(self):
        out = []
        for ctx in self.comm.ctxs:
            t = np.zeros(L.T_COUNT)
            L.check(self.comm.lib.graft_timers_get(ctx, L.ptr(t)))
            out.append(t)
        return out

    def stats(self):
        out = []
        for ctx in self.comm.ctxs:
            s = np.zeros(8, dtype=np.int64)
            L.check(self.comm.lib.graft_stats_get(ctx, L.ptr(s)))
            out.append(dict(ncells=int(s[0]), ncoo=int(s[1]), nnz=int(s[2]), nrows=int(s[3]), ncols=int(s[4]), launches=int(s[5]),
                            bytes_dev=int(s[6]), path={0: "none", 1: "unfused", 2: "fused-affine", 3: "fused-sweep", 4: "sumfact-gather"}[int(s[7])]))
        return out


def SparseMatrixAssembler(trial, test, strategy=None, **kw):
    """``SparseMatrixAssembler(trial,test[,strategy])`` (reference FESpaces.jl:853-862).  Keyword
    ``index_base`` plays the role of the ``Bi`` parameter of ``SparseMatrixCSR{Bi,Float64,Int}``
    (reference test/PoissonTests.jl:40)."""
    return GraftSparseMatrixAssembler(trial, test, strategy, **kw)


def _check_triangulation(form, assem):
    full = all(len(l) == m.num_cells for l, m in zip(form.dΩ.trian.cell_lids, assem.model.models))
    if isinstance(assem.strategy, FullyAssembledRows) and not full and assem.backend.nparts > 1:
        # same requirement as the reference: the user must integrate on Triangulation(FullyAssembledRows(),model)
        raise ValueError("FullyAssembledRows needs a triangulation with ghost cells (reference test/FESpacesTests.jl:171)")


def assemble_matrix_and_vector(form, assem):
    _check_triangulation(form, assem)
    assem._set_form(form)
    assem._symbolic(form)
    assem._numeric(3, form)
    return assem._wrap(), assem._vectors()


def assemble_matrix(form, assem):
    _check_triangulation(form, assem)
    assem._set_form(form)
    assem._symbolic(form)
    assem._numeric(1, form)
    return assem._wrap()


def assemble_vector(form, assem):
    _check_triangulation(form, assem)
    assem._set_form(form)
    assem._symbolic(form)
    assem._numeric(2, form)  # vector only: the matrix values of this assembler are left alone (the lifting evaluates what it needs)
    return assem._vectors()


def allocate_matrix_and_vector(form, assem):
    """Pattern only (Gridap ``allocate_matrix_and_vector``: the numeric loop runs with ``nothing``)."""
    _check_triangulation(form, assem)
    assem._set_form(form)
    assem._symbolic(form)
    return assem._wrap(), None


def assemble_matrix_and_vector_b(A, b, form, assem):
    """The reference's ``assemble_matrix_and_vector!(A,b,assem,data)``: values only, the
    sparsity, index sets and exchange plans are reused; bitwise repeatable."""
    assem._set_form(form)
    assem._symbolic(form)
    assem._numeric(3, form)
    mats = [A] if isinstance(A, PSparseMatrix) else [m for row in A for m in row]
    for m in mats:
        m.generation = assem._matrix_generation        # A is the matrix being re-assembled: its handle stays valid
    new = assem._vectors()
    if b is not None:                                  # like the reference's `!` variant the caller's b is updated in place
        olds = [b] if isinstance(b, PVector) else list(b)
        news = [new] if isinstance(new, PVector) else list(new)
        for o, n in zip(olds, news):
            for vo, vn in zip(o.vector_partition, n.vector_partition):
                vo[...] = vn
        return A, b
    return A, new


class BlockPVector:
    """``BlockPVector`` (reference BlockPartitionedArrays.jl:45-60): one ``PVector`` per field."""

    def __init__(self, blocks):
        self.blocks = list(blocks)


class BlockPMatrix:
    """``BlockPMatrix`` (reference BlockPartitionedArrays.jl:62-78): a grid of ``PSparseMatrix`` blocks."""

    def __init__(self, blocks):
        self.blocks = [list(r) for r in blocks]

    @property
    def blocksize(self):
        return len(self.blocks), len(self.blocks[0])


def _is_empty_block(A: PSparseMatrix):
    lib = A.assem.comm.lib
    for ctx in A.assem.comm.ctxs:
        m, n, nnz = L.c_i64(), L.c_i64(), L.c_i64()
        L.check(lib.graft_csr_query(ctx, A.bi, A.bj, C.byref(m), C.byref(n), C.byref(nnz)))
        if nnz.value:
            return False
    return True


def mul(c, A, b, alpha=1.0, beta=0.0):
    """``mul!(c,A,b,α,β)``: c_own <- β c_own + α (A b)_own ; b's ghost values are made consistent.
    Block arguments follow reference BlockPartitionedArrays.jl:336-351: y_i <- β y_i, then y_i += α A_ij x_j for every j."""
    if isinstance(A, BlockPMatrix):
        ni, nj = A.blocksize
        for i in range(ni):
            first = True
            for j in range(nj):
                if _is_empty_block(A.blocks[i][j]):  # a block the form does not couple is an empty matrix
                    continue
                mul(c.blocks[i], A.blocks[i][j], b.blocks[j], alpha, beta if first else 1.0)
                first = False
            if first:  # no coupled block in this block row
                for v in c.blocks[i].vector_partition:
                    v[:] = 0.0 if beta == 0.0 else beta * v
        return c
    assem = A.assem
    A._check_fresh()
    xs = [np.ascontiguousarray(v, dtype=np.float64) for v in b.vector_partition]
    ys = [np.ascontiguousarray(v[: ids.own_length], dtype=np.float64) for v, ids in zip(c.vector_partition, A.row_partition.indices)]
    for x, ids in zip(xs, A.col_partition.indices):
        assert len(x) == ids.local_length, "b must be laid out on the columns of A (change_ghost, reference Algebra.jl:270-293)"
    L.check(assem.comm.lib.graft_spmv(assem.comm.handle, A.bi, A.bj, float(alpha), L.ptr_array(xs), float(beta), L.ptr_array(ys)))
    for v, y, x, bx in zip(c.vector_partition, ys, xs, b.vector_partition):
        v[: len(y)] = y
        bx[:] = x
    return c


def cg(A: PSparseMatrix, b: PVector, x0=None, rtol=1e-12, maxit=1000, jacobi=True):
    """Jacobi-preconditioned CG for ``A x = b`` on the device (``graft_cg``): what a Krylov package does with
    ``mul!(y,A,x)``; the reference's tests use ``\\`` (gather + LU, test/FESpacesTests.jl:23).  Returns the solution as a
    PVector on the rows of ``A`` (own values; ghosts zero) and ``(iterations, relative residual)``."""
    assem = A.assem
    A._check_fresh()
    bs = [np.ascontiguousarray(v[: ids.own_length], dtype=np.float64) for v, ids in zip(b.vector_partition, A.row_partition.indices)]
    xs = [np.zeros(ids.own_length) if x0 is None else np.ascontiguousarray(v[: ids.own_length], dtype=np.float64)
          for v, ids in zip((x0.vector_partition if x0 is not None else bs), A.row_partition.indices)]
    it, rr = L.c_i32(), L.c_dbl()
    L.check(assem.comm.lib.graft_cg(assem.comm.handle, A.bi, L.ptr_array(bs), L.ptr_array(xs), float(rtol), int(maxit), int(bool(jacobi)),
                                    C.byref(it), C.byref(rr)))
    out = []
    for x, ids in zip(xs, A.row_partition.indices):
        v = np.zeros(ids.local_length)
        v[: ids.own_length] = x
        out.append(v)
    return PVector(out, A.row_partition), (it.value, rr.value)


def pvector_on_cols(A: PSparseMatrix, global_values):
    """A PVector on cols(A) holding ``global_values[gid-1]`` on own entries and zeros on ghosts."""
    vals = []
    for ids in A.col_partition.indices:
        v = np.zeros(ids.local_length)
        v[: ids.own_length] = global_values[ids.l2g[: ids.own_length] - 1]
        vals.append(v)
    return PVector(vals, A.col_partition)


def pvector_on_rows(A: PSparseMatrix, fill=0.0):
    return PVector([np.full(ids.local_length, fill) for ids in A.row_partition.indices], A.row_partition)
