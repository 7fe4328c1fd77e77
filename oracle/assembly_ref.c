/* assembly_ref.c -- CPU ORACLE (test infrastructure, NOT product code), plain C restatement of the reference's
 * single-part assembly path, used (a) to cross-check the numpy oracle (tests/test_oracle_c.py) and (b) as the
 * CPU baseline of bench.py (`cpu_baseline`, `--impl reference`): one OS thread per mesh part, like one MPI rank
 * per part in the reference's with_mpi mode.  Only tests/ and bench.py's baseline legs may load it.
 *
 * PARITY STATUS: "parity unpinned at entry level" -- see oracle/assembly_oracle.py: the reference is Julia, cannot
 * run in this image, and holds no golden matrices.  This file follows, step by step:
 *   A1  integrate(f, ::DistributedMeasure) -> Gridap.integrate per part    reference src/CellData.jl:234-239 [ext A.4]
 *         J_q = sum_v x_v (x) grad N_v(xi_q); grad phi_i = J_q^{-T} gradhat phi_i;
 *         K[i][j] = sum_q w_q |det J_q| grad phi_i . grad phi_j;  f[i] = sum_q w_q |det J_q| phi_i f
 *   A2  collect_cell_matrix_and_vector: Dirichlet lifting f -= K[:,D] g_D    reference src/FESpaces.jl:703-715 [ext A.7]
 *   A3  numeric_loop_matrix_and_vector! -> add_entries!: one COO triplet per (cell, li, lj), column-outer /
 *         row-inner, ids <= 0 skipped; b[i] += f[li]                          reference src/FESpaces.jl:794-798,
 *                                                                              src/Algebra.jl:811-821 [ext A.2, A.3]
 *   A9  create_from_nz -> sparse_from_coo -> sparsecsr: COO -> CSR, duplicates summed in (stable) sorted order
 *                                                                              reference src/Algebra.jl:574,636 [ext A.2]
 * Int64 indices and Float64 values like SparseMatrixCSR{0,Float64,Int} (reference test/PoissonTests.jl:40).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int64_t m, nnz, ncoo;
  int64_t* rowptr;
  int64_t* colind;
  double* vals;
  double* b;
} ref_csr;

static void gauss3(double* x, double* w) { /* 3-point Gauss-Legendre on [0,1]: Measure(Omega, 4) */
  double s = sqrt(0.6);
  x[0] = 0.5 * (1 - s); x[1] = 0.5; x[2] = 0.5 * (1 + s);
  w[0] = 5.0 / 18.0; w[1] = 8.0 / 18.0; w[2] = 5.0 / 18.0;
}
static void lagrange2(double x, double* v, double* d) { /* equispaced quadratic Lagrange basis, nodes 0, 1/2, 1 */
  v[0] = 2 * (x - 0.5) * (x - 1); v[1] = -4 * x * (x - 1); v[2] = 2 * x * (x - 0.5);
  d[0] = 4 * x - 3; d[1] = -8 * x + 4; d[2] = 4 * x - 1;
}

/* 3-D Q2 scalar Poisson on hexes.  cell_x: ncells x 8 x 3 vertex coordinates (x fastest); tix: 27 x 3 tensor index of
 * every local dof (from the reference element's node coordinates, never hard-coded); cell_dof_ids: ncells x 27
 * (1-based, <= 0 Dirichlet: value dirichlet_values[-id-1]); source: constant f. */
int ref_assemble_poisson_q2(int64_t ncells, const double* cell_x, const int32_t* tix, const int32_t* cell_dof_ids, int64_t nfree,
                            const double* dirichlet_values, double source, ref_csr* out) {
  enum { ND = 27, NQ = 27 };
  double gx[3], gw[3], val[3][3], der[3][3];
  gauss3(gx, gw);
  for (int q = 0; q < 3; ++q) lagrange2(gx[q], val[q], der[q]);
  /* reference tables: phi[q][i], dphi[q][i][a]; geometry Q1: dN[q][v][a] */
  static double phi[NQ][ND], dphi[NQ][ND][3], w[NQ], dN[NQ][8][3];
  for (int q = 0; q < NQ; ++q) {
    int qi[3] = {q % 3, (q / 3) % 3, q / 9};
    w[q] = gw[qi[0]] * gw[qi[1]] * gw[qi[2]];
    for (int i = 0; i < ND; ++i) {
      const int32_t* t = tix + 3 * i;
      phi[q][i] = val[qi[0]][t[0]] * val[qi[1]][t[1]] * val[qi[2]][t[2]];
      dphi[q][i][0] = der[qi[0]][t[0]] * val[qi[1]][t[1]] * val[qi[2]][t[2]];
      dphi[q][i][1] = val[qi[0]][t[0]] * der[qi[1]][t[1]] * val[qi[2]][t[2]];
      dphi[q][i][2] = val[qi[0]][t[0]] * val[qi[1]][t[1]] * der[qi[2]][t[2]];
    }
    for (int v = 0; v < 8; ++v) {
      double n1[3][2], d1[3][2];
      for (int a = 0; a < 3; ++a) { n1[a][0] = 1 - gx[qi[a]]; n1[a][1] = gx[qi[a]]; d1[a][0] = -1; d1[a][1] = 1; }
      int b0 = v & 1, b1 = (v >> 1) & 1, b2 = (v >> 2) & 1;
      dN[q][v][0] = d1[0][b0] * n1[1][b1] * n1[2][b2];
      dN[q][v][1] = n1[0][b0] * d1[1][b1] * n1[2][b2];
      dN[q][v][2] = n1[0][b0] * n1[1][b1] * d1[2][b2];
    }
  }
  /* A3s symbolic loop: count the triplets (nz_counter) */
  int64_t ncoo = 0;
  for (int64_t c = 0; c < ncells; ++c) {
    int nf = 0;
    for (int l = 0; l < ND; ++l) nf += cell_dof_ids[c * ND + l] > 0;
    ncoo += (int64_t)nf * nf;
  }
  int64_t* I = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncoo ? ncoo : 1));
  int64_t* J = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncoo ? ncoo : 1));
  double* V = (double*)malloc(sizeof(double) * (size_t)(ncoo ? ncoo : 1));
  double* b = (double*)calloc((size_t)(nfree ? nfree : 1), sizeof(double));
  if (!I || !J || !V || !b) return -1;
  /* A1-A3 numeric loop */
  int64_t k = 0;
  for (int64_t c = 0; c < ncells; ++c) {
    const double* X = cell_x + c * 24;
    const int32_t* ids = cell_dof_ids + c * ND;
    double K[ND][ND], F[ND], g[NQ][ND][3], wd[NQ];
    memset(K, 0, sizeof(K));
    memset(F, 0, sizeof(F));
    for (int q = 0; q < NQ; ++q) {
      double Jm[3][3] = {{0}};
      for (int v = 0; v < 8; ++v)
        for (int d = 0; d < 3; ++d)
          for (int a = 0; a < 3; ++a) Jm[d][a] += X[v * 3 + d] * dN[q][v][a];
      double c00 = Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1], c01 = Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2],
             c02 = Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0];
      double det = Jm[0][0] * c00 + Jm[0][1] * c01 + Jm[0][2] * c02, id = 1.0 / det;
      double Ji[3][3]; /* Ji[a][d] = d xi_a / d x_d */
      Ji[0][0] = c00 * id; Ji[1][0] = c01 * id; Ji[2][0] = c02 * id;
      Ji[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) * id;
      Ji[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) * id;
      Ji[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) * id;
      Ji[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) * id;
      Ji[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) * id;
      Ji[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) * id;
      wd[q] = w[q] * fabs(det);
      for (int i = 0; i < ND; ++i)
        for (int d = 0; d < 3; ++d) g[q][i][d] = dphi[q][i][0] * Ji[0][d] + dphi[q][i][1] * Ji[1][d] + dphi[q][i][2] * Ji[2][d];
    }
    for (int i = 0; i < ND; ++i)
      for (int j = 0; j < ND; ++j) {
        double acc = 0.0;
        for (int q = 0; q < NQ; ++q) acc += wd[q] * (g[q][i][0] * g[q][j][0] + g[q][i][1] * g[q][j][1] + g[q][i][2] * g[q][j][2]);
        K[i][j] = acc;
      }
    for (int i = 0; i < ND; ++i) {
      double acc = 0.0;
      for (int q = 0; q < NQ; ++q) acc += wd[q] * phi[q][i];
      F[i] = acc * source;
    }
    /* A2 lifting */
    for (int j = 0; j < ND; ++j)
      if (ids[j] <= 0) {
        double gd = dirichlet_values[-ids[j] - 1];
        for (int i = 0; i < ND; ++i) F[i] -= K[i][j] * gd;
      }
    /* A3 add_entries!: column outer, row inner */
    for (int lj = 0; lj < ND; ++lj) {
      if (ids[lj] <= 0) continue;
      for (int li = 0; li < ND; ++li) {
        if (ids[li] <= 0) continue;
        I[k] = ids[li]; J[k] = ids[lj]; V[k] = K[li][lj]; ++k;
      }
    }
    for (int li = 0; li < ND; ++li)
      if (ids[li] > 0) b[ids[li] - 1] += F[li];
  }
  /* A9 COO -> CSR: stable counting sort by row, then per row a stable sort by column + duplicate sum */
  int64_t m = nfree;
  int64_t* cnt = (int64_t*)calloc((size_t)m + 1, sizeof(int64_t));
  for (int64_t t = 0; t < ncoo; ++t) cnt[I[t]]++;
  for (int64_t r = 0; r < m; ++r) cnt[r + 1] += cnt[r];
  int64_t* bj = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncoo ? ncoo : 1));
  double* bv = (double*)malloc(sizeof(double) * (size_t)(ncoo ? ncoo : 1));
  int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * ((size_t)m + 1));
  memcpy(pos, cnt, sizeof(int64_t) * ((size_t)m + 1));
  for (int64_t t = 0; t < ncoo; ++t) { int64_t p = pos[I[t] - 1]++; bj[p] = J[t] - 1; bv[p] = V[t]; }
  free(I); free(J); free(V);
  int64_t* rowptr = (int64_t*)malloc(sizeof(int64_t) * ((size_t)m + 1));
  int64_t nnz = 0;
  rowptr[0] = 0;
  for (int64_t r = 0; r < m; ++r) {
    int64_t s = cnt[r], e = cnt[r + 1];
    for (int64_t a = s + 1; a < e; ++a) { /* stable insertion sort by column */
      int64_t cj = bj[a]; double cv = bv[a]; int64_t p = a - 1;
      while (p >= s && bj[p] > cj) { bj[p + 1] = bj[p]; bv[p + 1] = bv[p]; --p; }
      bj[p + 1] = cj; bv[p + 1] = cv;
    }
    int64_t o = nnz;
    for (int64_t a = s; a < e; ++a) {
      if (a > s && bj[a] == bj[nnz - 1] && nnz > o) bv[nnz - 1] += bv[a];
      else { bj[nnz] = bj[a]; bv[nnz] = bv[a]; ++nnz; }
    }
    rowptr[r + 1] = nnz;
  }
  free(cnt); free(pos);
  out->m = m; out->nnz = nnz; out->ncoo = ncoo; out->rowptr = rowptr; out->colind = bj; out->vals = bv; out->b = b;
  return 0;
}

void ref_csr_free(ref_csr* c) {
  free(c->rowptr); free(c->colind); free(c->vals); free(c->b);
  memset(c, 0, sizeof(*c));
}
