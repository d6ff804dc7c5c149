"""CPU study (test infrastructure, uses the oracle's assembled K): at EQUAL coarse size, rigid-body modes (6 per aggregate) against
affine modes (12 per aggregate: u = t + A rho, i.e. the rigid-body modes plus the six constant-strain modes) on half as many
aggregates -- additive two-level M^-1 = D^-1 + Z E^-1 Z^T, PCG iterations of the bench's Newton-step system.

    python tests/studies/affine_modes_study.py [cells]
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
    print(f"cells={cells} n={n} {name:44s} iters={it[0]}", flush=True)


def rcb(ids, parts, first, out):
    if parts <= 1:
        out[ids] = first
        return
    x = mesh.xyz[ids]
    ax = int(np.argmax(x.max(axis=0) - x.min(axis=0)))
    p1 = parts // 2
    mid = len(ids) * p1 // parts
    order = np.lexsort((ids, x[:, ax]))
    rcb(ids[order[:mid]], p1, first, out)
    rcb(ids[order[mid:]], parts - p1, first + p1, out)


def coarse(nagg, modes):
    aid = np.zeros(mesh.xyz.shape[0], int)
    rcb(np.arange(mesh.xyz.shape[0]), nagg, 0, aid)
    cen = np.stack([np.bincount(aid, weights=mesh.xyz[:, c], minlength=nagg) / np.bincount(aid, minlength=nagg) for c in range(3)], axis=1)
    k = {"t": 3, "rbm": 6, "affine": 12}[modes]
    rows, cols, vals = [], [], []
    nd, c = free // 3, free % 3
    rho = mesh.xyz[nd] - cen[aid[nd]]
    fi = np.arange(n)
    rows.append(fi); cols.append(k * aid[nd] + c); vals.append(np.ones(n))
    if modes == "rbm":
        for j in range(3):
            w = np.zeros(3); w[j] = 1.0
            u = np.cross(w, rho)[fi, c]
            rows.append(fi); cols.append(k * aid[nd] + 3 + j); vals.append(u)
    if modes == "affine":  # u_c = sum_j A[c, j] rho_j: mode (c, j) moves component c by rho_j
        for j in range(3):
            rows.append(fi); cols.append(k * aid[nd] + 3 + 3 * c + j); vals.append(rho[:, j])
    Z = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, k * nagg))
    E = (Z.T @ K @ Z).toarray()
    E += 1e-9 * np.diag(np.diag(E))
    Einv = np.linalg.inv(E)
    run(spla.LinearOperator((n, n), lambda r: r / d + Z @ (Einv @ (Z.T @ r))), f"jacobi + {modes} on {nagg} aggregates (nc = {k * nagg})")


run(spla.LinearOperator((n, n), lambda r: r / d), "jacobi")
for nagg, modes in ((64, "rbm"), (32, "affine"), (128, "t"), (128, "rbm"), (64, "affine"), (256, "rbm"), (128, "affine")):
    coarse(nagg, modes)
