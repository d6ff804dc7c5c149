"""End-to-end known answers of the reference's example problems, reproduced by the oracle's Newton driver
(direct sparse solve) -- the analytic checks are the reference's own."""
import numpy as np
import pytest

from onsas_jl_b200 import meshgen as mg
from tests import cases
from tests.golden import reference_vectors as G


def _uniaxial(oracle, mat, p, nsteps, tol, max_iter, grid=(1, 1, 1), linear="direct", **kw):
    m, mesh = cases.box_model(*grid, mat=mat)
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (1.0, 0.0, 0.0))
    lf = np.linspace(1.0 / nsteps, 1.0, nsteps)
    res = oracle.newton_solve(m, lf, lambda t: unit * (p * t), oracle.ConvergenceSettings(tol, tol, max_iter),
                              linear=linear, **kw)
    corner = int(np.argmax(mesh.xyz @ np.ones(3)))  # node at (Lx, Ly, Lz)
    U = res.U[-1].reshape(-1, 3)
    alpha = 1 + U[corner, 0] / 2.0
    beta = 1 + U[corner, 1] / 1.0
    gamma = 1 + U[corner, 2] / 1.0
    return m, mesh, res, alpha, beta, gamma


def test_uniaxial_extension_case1(oracle):
    """examples/uniaxial_extension/uniaxial_extension.jl:11-24,116-124,179-205: E=1, nu=0.3, p=3, 8 steps, tol 1e-8."""
    m, mesh, res, alpha, beta, gamma = _uniaxial(oracle, "svk", 3.0, 8, 1e-8, 30)
    E, nu, p = 1.0, 0.3, 3.0
    assert alpha == pytest.approx(2.0, rel=1e-4)            # root of E/2 a (a^2-1) = p
    assert beta == pytest.approx(np.sqrt(0.1), rel=1e-4)    # sqrt(1 - nu (a^2-1))
    assert gamma == pytest.approx(beta, rel=1e-10)
    assert res.iterations == G.UNIAXIAL_EXTENSION_ITERS
    # load factor from the displacement (load_factors_analytic :147-150) for every step
    for t, U in zip(np.linspace(1 / 8, 1, 8), res.U):
        ux = U.reshape(-1, 3)[int(np.argmax(mesh.xyz @ np.ones(3))), 0]
        assert (1 / p * E * 0.5 * ((1 + ux / 2) ** 3 - (1 + ux / 2))) == pytest.approx(t, rel=1e-4)
    # C and P of any element (homogeneous state) at the last ASSEMBLED configuration
    F = np.diag([alpha, beta, beta])
    Cc = F.T @ F
    lam, Gs = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    Eg = 0.5 * (Cc - np.eye(3))
    P = F @ (lam * np.trace(Eg) * np.eye(3) + 2 * Gs * Eg)
    np.testing.assert_allclose(res.tet_eps[-1][3].reshape(3, 3, order="F"), Cc, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(res.tet_sig[-1][3].reshape(3, 3, order="F"), P, rtol=1e-4, atol=1e-8)
    # reactions: sum of F_int on the x = 0 face balances the applied load p * Ly * Lz
    assert res.F_int[-1].reshape(-1, 3)[mesh.node_sets["x0"], 0].sum() == pytest.approx(-3.0, rel=1e-8)


def test_uniaxial_extension_refined_mesh_same_answer(oracle):
    """The homogeneous deformation is exact on every tet mesh of the box."""
    _, mesh, res, alpha, beta, _ = _uniaxial(oracle, "svk", 3.0, 8, 1e-8, 30, grid=(4, 2, 3))
    assert alpha == pytest.approx(2.0, rel=1e-6) and beta == pytest.approx(np.sqrt(0.1), rel=1e-6)
    np.testing.assert_allclose(res.U[-1], mg.homogeneous_field(mesh.xyz, alpha, beta), atol=1e-7)
    assert res.iterations == G.UNIAXIAL_EXTENSION_ITERS


def test_uniaxial_compression_case1(oracle):
    """examples/uniaxial_compression/uniaxial_compression.jl:11-25,177-181,239-293: NeoHookean, p=1, 9 steps, tol 1e-10."""
    m, mesh, res, alpha, beta, gamma = _uniaxial(oracle, "neo", -1.0, 9, 1e-10, 20)
    mu, K = 1.0 / 2.6, 1.0 / (3 * 0.4)
    assert alpha == pytest.approx(G.UNIAXIAL_COMPRESSION_ALPHA, rel=1e-9)
    assert beta == pytest.approx(G.UNIAXIAL_COMPRESSION_BETA, rel=1e-9)
    P11 = mu * alpha - mu / alpha + K * beta ** 2 * (alpha * beta ** 2 - 1)
    P22 = mu * beta - mu / beta + K * beta * (alpha ** 2 * beta ** 2 - alpha)
    assert P11 == pytest.approx(-1.0, rel=1e-4) and abs(P22) < 1e-8
    Pnum = res.tet_sig[-1][0].reshape(3, 3, order="F")
    assert Pnum[0, 0] == pytest.approx(P11, rel=1e-4) and abs(Pnum[1, 1]) < 1e-8 and abs(Pnum[2, 2]) < 1e-8
    assert res.iterations == G.UNIAXIAL_COMPRESSION_ITERS


@pytest.mark.parametrize("name,strain", [("roteng", 0), ("green", 1)])
def test_von_mises_truss(oracle, name, strain):
    """examples/von_misses_truss/von_misses_truss.jl:110-128: analytic load-displacement, both strain models."""
    m, fext, p = cases.von_mises_truss(strain)
    res = oracle.newton_solve(m, np.linspace(0.2, 1.0, 5), fext, oracle.ConvergenceSettings(1e-10, 1e-10, 10))
    E, A, H, V, Lr, Fk = (p[k] for k in ("E", "A", "H", "V", "L", "Fk"))
    for t, U in zip(np.linspace(0.2, 1.0, 5), res.U):
        uk = U[5]
        assert abs(U[3]) <= 100 * np.finfo(float).eps and abs(U[4]) <= np.finfo(float).eps
        if strain == 0:
            lam = -2 * E * A * ((H + uk) ** 2 + V ** 2 - Lr ** 2) / (Lr * (Lr + np.sqrt((H + uk) ** 2 + V ** 2))) * (H + uk) / np.sqrt((H + uk) ** 2 + V ** 2)
        else:
            lam = -2 * E * A * ((H + uk) * (2 * H * uk + uk ** 2)) / (2.0 * Lr ** 3)
        assert lam == pytest.approx(-t * Fk, rel=1e-4)
    assert res.U[-1][5] == pytest.approx(G.VON_MISES_UK[name], rel=1e-9)
    assert res.iterations == [G.VON_MISES_ITERS[name]] * 5


def test_clamped_truss(oracle):
    """examples/clamped_truss/clamped_truss.jl:85-98: tip force vs Green-strain analytic, rtol 1e-3; default tolerances."""
    m, fext, p = cases.clamped_truss(100)
    lf = np.linspace(0.1, 1.0, 10)
    res = oracle.newton_solve(m, lf, fext, oracle.ConvergenceSettings())
    for t, U in zip(lf, res.U):
        u = U[-1]
        eg = 0.5 * ((p["L"] + u) ** 2 - p["L"] ** 2) / p["L"] ** 2
        assert (p["L"] + u) / p["L"] * p["E"] * eg * p["A"] == pytest.approx(p["F"] * t, rel=1e-3)


def test_cylinder_lame_linear(oracle):
    """examples/cylinder_internal_pressure/cylinder_internal_pressure.jl:203-214: u_r = A r + B/r (plane strain),
    IsotropicLinearElastic, one Newton iteration from U = 0 is the linear solve.  atol = 1e-2 (Re - Ri)."""
    Ri, Re, Lz, E, nu, p = 100.0, 200.0, 30.0, 210.0, 0.3, 10.0
    mesh = mg.cylinder_tet_mesh(6, 32, 2, Ri, Re, Lz)
    fixed = {2: mesh.node_sets["z_caps"], 0: mesh.node_sets["outer_on_y_axis"], 1: mesh.node_sets["outer_on_x_axis"]}
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, fixed)
    m = oracle.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[oracle.MAT_ISOLINEAR], mat_params=[[E, nu]], free_dofs=free)
    Fp = mg.pressure_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["inner"], p)
    # total radial load = p * inner area
    r = np.linalg.norm(mesh.xyz[:, :2], axis=1)
    er = mesh.xyz[:, :2] / r[:, None]
    assert (Fp.reshape(-1, 3)[:, :2] * er).sum() == pytest.approx(p * 2 * np.pi * Ri * Lz, rel=1e-2)
    res = oracle.newton_solve(m, [1.0], lambda t: Fp * t, oracle.ConvergenceSettings(1e-8, 1e-8, 5))
    U = res.U[-1].reshape(-1, 3)
    ur = (U[:, :2] * er).sum(axis=1)
    A = (1 + nu) * (1 - 2 * nu) * Ri ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    B = (1 + nu) * Ri ** 2 * Re ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    np.testing.assert_allclose(ur, A * r + B / r, atol=1e-2 * (Re - Ri))
    assert np.abs(ur - (A * r + B / r)).max() < 0.05 * np.abs(ur).max()   # much tighter than the reference's atol
    assert np.abs(U[:, 2]).max() < 1e-2 * np.abs(ur).max() + 1e-9          # plane strain between the fixed caps
    assert res.iterations[0] <= 2


def test_newton_with_cg_matches_direct(oracle):
    """Reference default linear solve (un-preconditioned CG, reltol sqrt(eps)) and Jacobi-PCG with a tight
    tolerance both reproduce the direct-solve displacements within the north star's 1e-8."""
    args = ("svk", 3.0, 4, 1e-8, 30)
    _, _, rd, a_d, b_d, _ = _uniaxial(oracle, *args, grid=(3, 2, 2))
    _, _, rc, a_c, b_c, _ = _uniaxial(oracle, *args, grid=(3, 2, 2), linear="cg", cg_reltol=1e-12, jacobi=True)
    assert cases.rel_err(rc.U[-1], rd.U[-1]) < 1e-8
    assert rc.iterations == rd.iterations
    _, _, rr, _, _, _ = _uniaxial(oracle, *args, grid=(3, 2, 2), linear="cg")
    assert cases.rel_err(rr.U[-1], rd.U[-1]) < 1e-6


def test_two_level_preconditioner_definition(oracle):
    """The preconditioner the streamed CG applies for precond = 2, restated with numpy on the
    oracle's K: M^-1 = D^-1 + Z E^-1 Z^T, Z = rigid-body modes (3 translations + 3 rotations about the centroid) of node
    aggregates restricted to the free dofs, E = Z^T K Z.  It has no counterpart in the reference, so what is pinned here
    is its definition: M^-1 is symmetric positive definite on the free dofs, PCG with it reaches the direct solution,
    and it needs fewer iterations than Jacobi and than the translations-only coarse space on the same aggregates."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m, mesh = cases.box_model(12, 6, 6, mat="neo", jitter=0.1)
    U = cases.random_U(m, 0.002)
    asm = oracle.Assembly(m).assemble(U)
    free = np.asarray(m.free_dofs)
    K = asm.csr()[free][:, free].tocsr()
    n = K.shape[0]
    d = K.diagonal()
    # aggregates: a 3 x 2 x 2 grid of boxes over the 2 x 1 x 1 domain
    g = np.minimum((mesh.xyz * [1.5, 2.0, 2.0]).astype(int), [2, 1, 1])
    agg = g[:, 0] + 3 * (g[:, 1] + 2 * g[:, 2])
    n_agg = int(agg.max()) + 1
    cen = np.stack([np.bincount(agg, weights=m.xyz[:, c], minlength=n_agg) for c in range(3)], axis=1) / np.bincount(agg)[:, None]

    def coarse_basis(rotations):
        cd = 6 if rotations else 3
        rows, cols, vals = [], [], []
        for fi, dof in enumerate(free):
            nd, c = dof // 3, dof % 3
            rows.append(fi); cols.append(cd * agg[nd] + c); vals.append(1.0)
            if rotations:
                rho = m.xyz[nd] - cen[agg[nd]]
                for k in range(3):                      # u = e_k x rho
                    u = np.cross(np.eye(3)[k], rho)
                    rows.append(fi); cols.append(cd * agg[nd] + 3 + k); vals.append(u[c])
        return sp.csr_matrix((vals, (rows, cols)), shape=(n, cd * n_agg))

    b = np.random.default_rng(8).standard_normal(n)
    xd = spla.spsolve(K.tocsc(), b)

    def pcg_iters(Minv):
        it = [0]
        x, info = spla.cg(K, b, rtol=1e-10, atol=0.0, maxiter=5000, M=Minv, callback=lambda _: it.__setitem__(0, it[0] + 1))
        assert info == 0 and np.abs(x - xd).max() / np.abs(xd).max() < 1e-7
        return it[0]

    its = {"jacobi": pcg_iters(spla.LinearOperator((n, n), lambda r: r / d))}
    for name, rot in (("translations", False), ("rigid_body", True)):
        Z = coarse_basis(rot)
        E = (Z.T @ K @ Z).toarray()
        assert np.allclose(E, E.T, rtol=1e-12, atol=1e-14) and np.linalg.eigvalsh(E).min() > 0
        Einv = np.linalg.inv(E)
        Minv = lambda r, Z=Z, Einv=Einv: r / d + Z @ (Einv @ (Z.T @ r))  # noqa: E731
        # symmetric positive definite: x' M^-1 y = y' M^-1 x and x' M^-1 x > 0 on random vectors
        rng = np.random.default_rng(1)
        x1, y1 = rng.standard_normal(n), rng.standard_normal(n)
        assert x1 @ Minv(y1) == pytest.approx(y1 @ Minv(x1), rel=1e-10) and x1 @ Minv(x1) > 0
        its[name] = pcg_iters(spla.LinearOperator((n, n), Minv))
    assert its["rigid_body"] < its["translations"] < its["jacobi"], its
