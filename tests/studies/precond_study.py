"""CPU study (test infrastructure, uses the oracle's assembled K): PCG iteration counts of the Newton step's linear
system for candidate preconditioners -- Jacobi (shipped), 3x3 block-Jacobi, Jacobi + an aggregated coarse space
(piecewise-constant translations, optionally rigid-body rotations).  Numbers quoted in DESIGN.md section 9.

    python tests/studies/precond_study.py [cells]
"""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import bench
from oracle import oracle as O
from onsas_jl_b200 import meshgen as mg

def study(cells):
    mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
    m = O.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[O.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
    asm = O.AssemblyMT(m).assemble(U_prev)
    K = asm.csr()[free][:, free].tocsr()
    b = (Fext - asm.F_int)[free]
    n = K.shape[0]
    tol = 1.4901161193847656e-08
    def run(M, name):
        it = [0]
        def cb(xk): it[0] += 1
        x, info = spla.cg(K, b, rtol=tol, atol=0, maxiter=20000, M=M, callback=cb)
        print(f"cells={cells} n={n} {name:28s} iters={it[0]} info={info}", flush=True)
        return it[0]
    d = K.diagonal()
    run(spla.LinearOperator((n, n), lambda r: r / d), "jacobi")
    # block Jacobi 3x3 on nodes (free dofs subset: build per-node blocks over free dofs)
    node = free // 3
    # block-diagonal by node
    order = np.arange(n)
    Kc = K.tocoo()
    same = node[Kc.row] == node[Kc.col]
    B = sp.csr_matrix((Kc.data[same], (Kc.row[same], Kc.col[same])), shape=(n, n)).tocsc()
    Binv = spla.splu(B)
    run(spla.LinearOperator((n, n), lambda r: Binv.solve(r)), "block-jacobi 3x3")
    # two-level: aggregates of a^3 nodes, translations only / + rotations
    g = np.rint(mesh.xyz * cells).astype(int)
    for a, rbm in ((4, False), (4, True), (7, False), (7, True)):
        ag = (g // a)
        na = ag.max(axis=0) + 1
        aid = ag[:, 0] + na[0] * (ag[:, 1] + na[1] * ag[:, 2])
        nagg = int(aid.max()) + 1
        # Z: n_free x (3 or 6)*nagg
        rows, cols, vals = [], [], []
        cen = np.zeros((nagg, 3)); cnt = np.bincount(aid, minlength=nagg)
        for c in range(3): cen[:, c] = np.bincount(aid, weights=mesh.xyz[:, c], minlength=nagg) / np.maximum(cnt, 1)
        k = 6 if rbm else 3
        for fi, dof in enumerate(free):
            nd, c = dof // 3, dof % 3
            a_ = aid[nd]
            rows.append(fi); cols.append(k * a_ + c); vals.append(1.0)
            if rbm:
                x = mesh.xyz[nd] - cen[a_]
                # rotation modes: u = w x r ; w = e_j
                for j in range(3):
                    w = np.zeros(3); w[j] = 1.0
                    u = np.cross(w, x)
                    if u[c] != 0.0:
                        rows.append(fi); cols.append(k * a_ + 3 + j); vals.append(u[c])
        Z = sp.csr_matrix((vals, (rows, cols)), shape=(n, k * nagg))
        keep = np.asarray(abs(Z).sum(axis=0)).ravel() > 0
        Z = Z[:, keep]
        E = (Z.T @ K @ Z).toarray()
        E += 1e-12 * np.trace(E) / E.shape[0] * np.eye(E.shape[0])
        Einv = np.linalg.inv(E)
        def M2(r, Z=Z, Einv=Einv):
            return r / d + Z @ (Einv @ (Z.T @ r))
        run(spla.LinearOperator((n, n), M2), f"jacobi + coarse a={a} rbm={int(rbm)} nc={E.shape[0]}")
        # deflation-style multiplicative (A-DEF2-like): z = (I - Q A) D^-1 r + Q r  with Q = Z E^-1 Z^T
        def M3(r, Z=Z, Einv=Einv):
            z = r / d
            return z - Z @ (Einv @ (Z.T @ (K @ z))) + Z @ (Einv @ (Z.T @ r))
        run(spla.LinearOperator((n, n), M3), f"adef2 a={a} rbm={int(rbm)} nc={E.shape[0]}")

study(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
