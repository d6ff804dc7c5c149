"""CPU study (test infrastructure, uses the oracle's assembled K): PCG iteration counts of the Newton step's linear system for
an additive MULTILEVEL preconditioner on nested rigid-body-mode aggregates,
    M^-1 = D^-1 + sum_l Z_l B_l Z_l^T,   B_l = blockdiag(Z_l^T K Z_l)^-1 on the intermediate levels, (Z_L^T K Z_L)^-1 at the top,
against Jacobi and the shipped two-level method (one level of aggregates, exact coarse solve).
    python tests/studies/multilevel_study.py [cells] [mat]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
import bench
from oracle import oracle as O
from onsas_jl_b200 import meshgen as mg


def rbm_Z(xyz, free, aid, nagg):
    n = len(free)
    cen = np.zeros((nagg, 3)); cnt = np.bincount(aid, minlength=nagg)
    for c in range(3): cen[:, c] = np.bincount(aid, weights=xyz[:, c], minlength=nagg) / np.maximum(cnt, 1)
    nd, c = free // 3, free % 3
    a = aid[nd]
    rho = xyz[nd] - cen[a]
    rows = [np.arange(n)]; cols = [6 * a + c]; vals = [np.ones(n)]
    for j in range(3):   # u = e_j x rho
        w = np.zeros(3); w[j] = 1
        u = np.cross(w, rho)[np.arange(n), c]
        rows.append(np.arange(n)); cols.append(6 * a + 3 + j); vals.append(u)
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, 6 * nagg))


def study(cells, mat):
    if mat == "neo":
        mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
        m = O.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[O.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
    else:
        mesh = mg.box_tet_mesh(cells, cells, cells, 1.0, 1.0, 1.0)
        free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
        a0, b0 = bench.uniaxial_state(bench._P_svk, 3 * 7 / 8, (1.8, 0.5))
        U_prev = mg.homogeneous_field(mesh.xyz, a0, b0)
        Fext = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (3.0, 0, 0))
        m = O.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[O.MAT_SVK], mat_params=[[bench.LAM, bench.MU]], free_dofs=free)
    asm = O.AssemblyMT(m).assemble(U_prev)
    K = asm.csr()[free][:, free].tocsr()
    b = (Fext - asm.F_int)[free]
    n = K.shape[0]
    tol = 1.4901161193847656e-08
    d = K.diagonal()

    def run(M, name):
        it = [0]
        def cb(xk): it[0] += 1
        x, info = spla.cg(K, b, rtol=tol, atol=0, maxiter=20000, M=M, callback=cb)
        print(f"cells={cells} {mat} n={n} {name:60s} iters={it[0]}", flush=True)
        return it[0]

    run(spla.LinearOperator((n, n), lambda r: r / d), "jacobi")
    g = np.rint(mesh.xyz * cells / mesh.xyz.max(axis=0)).astype(int)

    def level(a):
        ag = g // a
        na = ag.max(axis=0) + 1
        aid = ag[:, 0] + na[0] * (ag[:, 1] + na[1] * ag[:, 2])
        nagg = int(aid.max()) + 1
        Z = rbm_Z(mesh.xyz, free, aid, nagg)
        E = (Z.T @ K @ Z).tocsr()
        return Z, E, nagg

    def exact_inv(E):
        Ed = E.toarray()
        dd = np.diag(Ed).copy(); dd[dd == 0] = 1.0
        Ed[np.diag_indices_from(Ed)] = np.where(np.diag(Ed) == 0, 1.0, np.diag(Ed) * (1 + 1e-9))
        return np.linalg.inv(Ed)

    def blockdiag_inv(E, nagg):
        Ed = E.tolil()
        blocks = []
        Ec = E.tocsr()
        out = sp.lil_matrix(E.shape)
        for a in range(nagg):
            B = Ec[6 * a:6 * a + 6, 6 * a:6 * a + 6].toarray()
            dz = np.diag(B) == 0
            B[dz, dz] = 1.0
            B[np.diag_indices(6)] *= (1 + 1e-9)
            out[6 * a:6 * a + 6, 6 * a:6 * a + 6] = np.linalg.inv(B)
        return out.tocsr()

    cfgs = [[int(x) for x in c.split(',')] for c in sys.argv[3].split(';')] if len(sys.argv) > 3 else ([8], [4], [4, 8], [4, 16], [2, 4, 8], [4, 8, 16], [3, 9], [2, 8])
    for sizes in cfgs:
        sizes = [s for s in sizes if s < cells]
        if not sizes: continue
        lv = [level(a) for a in sizes]
        ops = []
        for k, (Z, E, nagg) in enumerate(lv):
            if k == len(lv) - 1:
                Ei = exact_inv(E); ops.append((Z, Ei))
            else:
                ops.append((Z, blockdiag_inv(E, nagg)))
        for w in (1.0,):
            def M(r, ops=ops, w=w):
                z = r / d
                for k, (Z, B) in enumerate(ops):
                    wk = 1.0 if k == len(ops) - 1 else w
                    z = z + wk * (Z @ (B @ (Z.T @ r)))
                return z
            run(spla.LinearOperator((n, n), M), f"additive levels a={sizes} nc_top={6 * lv[-1][2]} w_mid={w}")
        # exact two-level on the FINEST aggregates (what the multilevel scheme approximates)
        if len(lv) > 1:
            Z, E, nagg = lv[0]
            if 6 * nagg <= 8000:
                Ei = exact_inv(E)
                run(spla.LinearOperator((n, n), lambda r, Z=Z, Ei=Ei: r / d + Z @ (Ei @ (Z.T @ r))), f"exact two-level on a={sizes[0]} nc={6 * nagg}")


study(int(sys.argv[1]) if len(sys.argv) > 1 else 16, sys.argv[2] if len(sys.argv) > 2 else "svk")
