"""graft: B200-native distributed FE assembly + halo SpMV behind the GridapDistributed API.

Host-side mirror (Python; the reference's Julia toolchain is absent from this image) of the
reference interface for ONE path: ``SparseMatrixAssembler`` / ``assemble_matrix_and_vector`` /
``AffineFEOperator`` / ``mul!`` (reference FESpaces.jl:677-862, Algebra.jl:395-1257).  All
arithmetic runs in ``lib/libgraft.so`` (hand-written sm_100a CUDA behind a C ABI, include/graft.h).
There is no CPU fallback: without the library or a GPU the assembly calls raise.
"""
from .parrays import DebugBackend, DistBackend, with_debug, with_dist, PRange, LocalIndices, OwnAndGhostIndices
from .geometry import (CartesianDiscreteModel, Triangulation, Measure, SubAssembledRows,
                       FullyAssembledRows, NCube, Boundary, facet_points)
from .fespaces import ReferenceFE, lagrangian, FESpace, TestFESpace, TrialFESpace, generate_gids
from .assembly import (SparseMatrixAssembler, GraftSparseMatrixAssembler, assemble_matrix_and_vector, assemble_matrix,
                       assemble_vector, allocate_matrix_and_vector, assemble_matrix_and_vector_b, AffineFEOperator,
                       PSparseMatrix, PVector, BlockPMatrix, BlockPVector, mul, cg, pvector_on_cols, pvector_on_rows, Poisson, Mass, LinearElasticity, StokesTH, PLaplacian,
                       CellArrays, collect_cell_matrix_and_vector)
from .problems import build_problem, build_stokes_problem, vertex_perturbation, interior_vertex_mask
from . import libgraft
