"""numpy model of the sum-factorised Q2 grad-grad cell matrix (design check of the sweep kernel; not product code).

K[i,j] = sum_q sum_ab G^{ab}(q) dhat_a phi_i(q) dhat_b phi_j(q),   G = w |det J| J^{-1} J^{-T}   (symmetric 3x3 per point)
With u_d = (i_d, j_d) the index pairs per direction and X_t[u, q] = (L or L')_i(q) (L or L')_j(q), t = (row is derivative, col is derivative):
  stage 1  T1^{ab}[u1][q2,q3]    = sum_q1 X_{t(a,b,1)}[u1,q1] G^{ab}[q1,q2,q3]
  stage 2  T2^{g}[u1,u2][q3]     = sum_{(a,b): t(a,b,3)=g} sum_q2 X_{t(a,b,2)}[u2,q2] T1^{ab}[u1][q2,q3]
  stage 3  K[u1,u2,u3]           = sum_g sum_q3 X_g[u3,q3] T2^g[u1,u2][q3]
Only the 6 pairs u1 = (i1 <= j1) are computed; K[(j1,i1),tau u2,tau u3] = K[(i1,j1),u2,u3].
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graft_import
from oracle import assembly_oracle as orc

g = graft_import.load()


def tables(ref_nodes, quad_degree=4):
    n = (quad_degree + 2) // 2
    x1, w1 = orc.gauss_legendre_01(n)
    L, Lp = orc.lagrange_1d(2, x1)  # [q, node]
    L, Lp = L.T.copy(), Lp.T.copy()  # [node, q]
    # X[t][(i,j)][q], t = 2*(row derivative) + (col derivative)
    X = np.zeros((4, 9, n))
    for ti in range(2):
        for tj in range(2):
            A = Lp if ti else L
            B = Lp if tj else L
            for i in range(3):
                for j in range(3):
                    X[2 * ti + tj, 3 * i + j] = A[i] * B[j]
    tix = np.rint(np.asarray(ref_nodes) * 2).astype(int)  # local dof -> tensor index
    node_of = np.zeros((3, 3, 3), dtype=int)
    for l, (a, b, c) in enumerate(tix):
        node_of[a, b, c] = l
    return x1, w1, X, tix, node_of


def geometry_G(Xv, x1, w1):
    """G[a,b,q1,q2,q3] and wd[q1,q2,q3] of a trilinear cell (vertices x fastest)."""
    n = len(x1)
    G = np.zeros((3, 3, n, n, n))
    wd = np.zeros((n, n, n))
    for q3 in range(n):
        for q2 in range(n):
            for q1 in range(n):
                xi = np.array([[x1[q1], x1[q2], x1[q3]]])
                _, dN = orc.geometry_tables(3, xi)
                J = np.einsum("vd,va->da", Xv, dN[0])
                Ji = np.linalg.inv(J)  # [a,d]
                w = w1[q1] * w1[q2] * w1[q3] * abs(np.linalg.det(J))
                G[:, :, q1, q2, q3] = w * Ji @ Ji.T
                wd[q1, q2, q3] = w
    return G, wd


def typ(a, b, d):
    return 2 * (a == d) + (b == d)


def cell_matrix_sumfact(Xv, ref_nodes):
    x1, w1, X, tix, node_of = tables(ref_nodes)
    n = len(x1)
    G, _ = geometry_G(Xv, x1, w1)
    K = np.zeros((27, 27))
    nfma = 0
    for i1 in range(3):
        for j1 in range(i1, 3):  # 6 pairs
            u1 = 3 * i1 + j1
            T2 = np.zeros((4, 9, n))  # [g][u2][q3]
            for q3 in range(n):  # work item (u1, q3): stages 1 + 2 in registers
                T1 = np.zeros((3, 3, n))
                for a in range(3):
                    for b in range(3):
                        for q2 in range(n):
                            for q1 in range(n):
                                T1[a, b, q2] += X[typ(a, b, 0), u1, q1] * G[a, b, q1, q2, q3]
                                nfma += 1
                for a in range(3):
                    for b in range(3):
                        gt = typ(a, b, 2)
                        for u2 in range(9):
                            for q2 in range(n):
                                T2[gt, u2, q3] += X[typ(a, b, 1), u2, q2] * T1[a, b, q2]
                                nfma += 1
            for u2 in range(9):  # work item (u1, third of u2): stage 3
                i2, j2 = divmod(u2, 3)
                for u3 in range(9):
                    i3, j3 = divmod(u3, 3)
                    v = 0.0
                    for gt in range(4):
                        for q3 in range(n):
                            v += X[gt, u3, q3] * T2[gt, u2, q3]
                            nfma += 1
                    li, lj = node_of[i1, i2, i3], node_of[j1, j2, j3]
                    K[li, lj] = v
                    if i1 != j1:
                        K[lj, li] = v
    return K, nfma


if __name__ == "__main__":
    poly = g.NCube(3)
    ref_nodes = poly.q2_ref_nodes()
    rng = np.random.default_rng(0)
    Xv = np.array([[(v >> d) & 1 for d in range(3)] for v in range(8)], dtype=float) * np.array([0.3, 0.2, 0.25])
    Xv += rng.uniform(-0.03, 0.03, Xv.shape)
    K, nfma = cell_matrix_sumfact(Xv, ref_nodes)
    Kref, _ = orc.integrate_cells(("poisson",), Xv[None], ref_nodes, 2, 1, 4)
    err = np.abs(K - Kref[0]).max() / np.abs(Kref[0]).max()
    print("max rel err", err, "fma per cell", nfma)
    assert err < 1e-13
