"""CPU study (test infrastructure, uses the oracle's assembled K): what a THIRD level would buy.  Fine aggregates (a_f^3 nodes, rigid-body
modes) give a sparse coarse operator E1 = Z1^T K Z1 that is too large for a dense inverse at bench size; a second aggregation (groups of
fine aggregates, rigid-body transfer T) gives the small dense E2 = T^T E1 T.  PCG iterations of the bench's Newton-step system with
  two-level (shipped):   M^-1 = D^-1 + Z E^-1 Z^T                         on coarse aggregates
  exact fine two-level:  M^-1 = D^-1 + Z1 E1^-1 Z1^T                      (the bound of any three-level scheme)
  additive three-level:  M^-1 = D^-1 + Z1 (B1^-1 + T E2^-1 T^T) Z1^T      B1 = the 6 x 6 block diagonal of E1
  V-cycle on level 1:    E1^-1 ~ S^T-smoothed coarse correction: y = w B1^-1 r; y += T E2^-1 T^T (r - E1 y); y += w B1^-1 (r - E1 y)

    python tests/studies/three_level_study.py [cells] [a_fine] [group]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import bench
from oracle import oracle as O

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 33
a_f = int(sys.argv[2]) if len(sys.argv) > 2 else 4
grp = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
m = O.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[O.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
asm = O.AssemblyMT(m).assemble(U_prev)
K = asm.csr()[free][:, free].tocsr()
b = (Fext - asm.F_int)[free]
n = K.shape[0]
d = K.diagonal()
tol = 1.4901161193847656e-08


def run(M, name):
    it = [0]

    def cb(xk):
        it[0] += 1
    x, info = spla.cg(K, b, rtol=tol, atol=0, maxiter=20000, M=M, callback=cb)
    print(f"cells={cells} n={n} {name:64s} iters={it[0]}", flush=True)


g = np.rint(mesh.xyz * cells).astype(int)
nd, comp = free // 3, free % 3


def rbm_Z(aid, nagg, pts, rows_of):
    """Z [len(rows_of) x 6 nagg]: row k = dof (point rows_of-th entry, component) of a point in aggregate aid."""
    cen = np.stack([np.bincount(aid, weights=pts[:, c], minlength=nagg) / np.maximum(np.bincount(aid, minlength=nagg), 1) for c in range(3)], axis=1)
    return cen


def grid_agg(a):
    ag = g // a
    na = ag.max(axis=0) + 1
    return ag[:, 0] + na[0] * (ag[:, 1] + na[1] * ag[:, 2]), int(np.prod(na)), na


def build_Z(aid, nagg):
    cen = np.stack([np.bincount(aid, weights=mesh.xyz[:, c], minlength=nagg) / np.maximum(np.bincount(aid, minlength=nagg), 1) for c in range(3)], axis=1)
    rho = mesh.xyz[nd] - cen[aid[nd]]
    fi = np.arange(n)
    rows, cols, vals = [fi], [6 * aid[nd] + comp], [np.ones(n)]
    for j in range(3):
        w = np.zeros(3)
        w[j] = 1.0
        rows.append(fi)
        cols.append(6 * aid[nd] + 3 + j)
        vals.append(np.cross(w, rho)[fi, comp])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, 6 * nagg)), cen


run(spla.LinearOperator((n, n), lambda r: r / d), "jacobi")
# shipped two-level at a comparable coarse size
aid_c, nagg_c, _ = grid_agg(a_f * grp)
Zc, _ = build_Z(aid_c, nagg_c)
Ec = (Zc.T @ K @ Zc).toarray()
Ec += 1e-9 * np.diag(np.diag(Ec))
Ecinv = np.linalg.inv(Ec)
run(spla.LinearOperator((n, n), lambda r: r / d + Zc @ (Ecinv @ (Zc.T @ r))), f"two-level, {a_f * grp}^3-node aggregates (nc = {Ec.shape[0]})")
# fine level
aid_f, nagg_f, na_f = grid_agg(a_f)
Z1, cen_f = build_Z(aid_f, nagg_f)
E1 = (Z1.T @ K @ Z1).tocsr()
E1 = E1 + sp.diags(1e-9 * E1.diagonal())
E1lu = spla.splu(E1.tocsc())
run(spla.LinearOperator((n, n), lambda r: r / d + Z1 @ E1lu.solve(Z1.T @ r)), f"exact fine two-level, {a_f}^3-node aggregates (nc = {E1.shape[0]}, nnz = {E1.nnz})")
# second aggregation: groups of grp^3 fine aggregates, rigid-body transfer T (child translation = t + omega x dvec, child rotation = omega)
fa = np.arange(nagg_f)
fijk = np.stack([fa % na_f[0], (fa // na_f[0]) % na_f[1], fa // (na_f[0] * na_f[1])], axis=1)
pa = fijk // grp
npa = pa.max(axis=0) + 1
pid = pa[:, 0] + npa[0] * (pa[:, 1] + npa[1] * pa[:, 2])
nagg_2 = int(np.prod(npa))
cnt_f = np.bincount(aid_f, minlength=nagg_f).astype(float)
cen_2 = np.stack([np.bincount(pid, weights=cen_f[:, c] * cnt_f, minlength=nagg_2) / np.maximum(np.bincount(pid, weights=cnt_f, minlength=nagg_2), 1) for c in range(3)], axis=1)
dv = cen_f - cen_2[pid]
rows, cols, vals = [], [], []
for a in range(nagg_f):
    for c in range(3):
        rows += [6 * a + c, 6 * a + 3 + c]
        cols += [6 * pid[a] + c, 6 * pid[a] + 3 + c]
        vals += [1.0, 1.0]
    for j in range(3):
        w = np.zeros(3)
        w[j] = 1.0
        u = np.cross(w, dv[a])
        for c in range(3):
            if u[c] != 0.0:
                rows.append(6 * a + c)
                cols.append(6 * pid[a] + 3 + j)
                vals.append(u[c])
T = sp.csr_matrix((vals, (rows, cols)), shape=(6 * nagg_f, 6 * nagg_2))
E2 = (T.T @ E1 @ T).toarray()
E2 += 1e-9 * np.diag(np.diag(E2))
E2inv = np.linalg.inv(E2)
# block diagonal of E1
E1c = E1.tocoo()
same = (E1c.row // 6) == (E1c.col // 6)
B1 = sp.csr_matrix((E1c.data[same], (E1c.row[same], E1c.col[same])), shape=E1.shape).tocsc()
B1lu = spla.splu(B1)


def additive3(r):
    w1 = Z1.T @ r
    return r / d + Z1 @ (B1lu.solve(w1) + T @ (E2inv @ (T.T @ w1)))


run(spla.LinearOperator((n, n), additive3), f"additive three-level ({a_f}^3 fine, {grp}^3 groups: nc1 = {E1.shape[0]}, nc2 = {E2.shape[0]})")
for omega in (0.5, 0.7, 1.0):
    def vcycle(r, omega=omega):
        w1 = Z1.T @ r
        y = omega * B1lu.solve(w1)
        y = y + T @ (E2inv @ (T.T @ (w1 - E1 @ y)))
        y = y + omega * B1lu.solve(w1 - E1 @ y)
        return r / d + Z1 @ y
    run(spla.LinearOperator((n, n), vcycle), f"level-1 V(1,1) cycle, block-Jacobi omega = {omega}")
for sweeps in (2, 4):
    def cheb_like(r, sweeps=sweeps):
        # E1^-1 ~ `sweeps` PCG-free Richardson sweeps with the additive two-level operator on level 1 (symmetric: same operator each sweep)
        w1 = Z1.T @ r
        P1 = lambda v: 0.7 * B1lu.solve(v) + T @ (E2inv @ (T.T @ v))
        y = P1(w1)
        for _ in range(sweeps - 1):
            y = y + 0.5 * P1(w1 - E1 @ y)
        return r / d + Z1 @ y
    run(spla.LinearOperator((n, n), cheb_like), f"level-1 Richardson x{sweeps} with the additive level-1 operator (not symmetric: indicative only)")
