"""End-to-end parity: the reference's examples solved through the mirrored API on the GPU vs the oracle's
Newton driver with a direct solve (displacements, reactions, iteration counts within 1e-8) and vs the analytic answers."""
import math

import numpy as np
import pytest

from onsas_jl_b200 import meshgen as mg
from tests import cases
from tests.golden import reference_vectors as G
from tests.test_host_logic import _uniaxial_structure

pytestmark = pytest.mark.gpu
SOLVE_RTOL = 1e-8


def test_uniaxial_extension_as_shipped(ob, oracle):
    """configs[0]: examples/uniaxial_extension Case 1 through Structure / NonLinearStaticAnalysis / NewtonRaphson / solve."""
    s, n, t = _uniaxial_structure(ob)
    sa = ob.NonLinearStaticAnalysis(s, NSTEPS=8)
    nr = ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 30), cg_reltol=1e-13, cg_maxiter=500)
    sol = ob.solve(sa, nr)
    assert sa.current_step == 1 and sa._ctx is None           # solve() works on a deep copy
    ux, uy, uz = sol.displacements(n[6])
    alpha, beta = 1 + ux[-1] / 2.0, 1 + uy[-1] / 1.0
    assert alpha == pytest.approx(2.0, rel=1e-4) and beta == pytest.approx(math.sqrt(0.1), rel=1e-4)   # :201-202
    fl = s.flat
    fm = oracle.FlatModel(xyz=fl.xyz, tets=fl.tets, mat_kind=fl.mat_kind, mat_params=fl.mat_params, free_dofs=fl.free_dofs)
    ref = oracle.newton_solve(fm, sa.load_factors(), fl.fext, oracle.ConvergenceSettings(1e-8, 1e-8, 30))
    assert sol.iterations() == ref.iterations == G.UNIAXIAL_EXTENSION_ITERS
    for k in range(8):
        assert cases.rel_err(sol.U[k], ref.U[k]) < SOLVE_RTOL
        assert cases.rel_err(sol.F_int[k], ref.F_int[k]) < SOLVE_RTOL
        assert cases.rel_err(sol.tet_stress[k], ref.tet_sig[k]) < SOLVE_RTOL
    e = t[2]
    F = np.diag([alpha, beta, beta])
    np.testing.assert_allclose(sol.strain(e)[-1], F.T @ F, rtol=1e-4, atol=1e-8)       # :203
    assert sol.reactions()[-1].reshape(-1, 3)[:4, 0].sum() == pytest.approx(-3.0, rel=1e-7)
    assert all(isinstance(c, (ob.ResidualForceCriterion, ob.DeltaUCriterion)) for c in sol.criterion())


@pytest.mark.parametrize("precond", ["jacobi", "none", "two_level"])
def test_uniaxial_compression_neohookean(ob, oracle, precond):
    """configs[1] at oracle-sized refinement: NeoHookean cube compression, 9 steps, tol 1e-10."""
    m, mesh = cases.box_model(6, 3, 3, mat="neo")
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (-1.0, 0.0, 0.0))
    s = ob.Structure.from_arrays(m.xyz, tets=m.tets, materials=[ob.NeoHookean(E=1.0, nu=0.3)], free_dofs=m.free_dofs,
                                 fext=lambda t: unit * t)
    sa = ob.NonLinearStaticAnalysis(s, NSTEPS=9)
    sol = ob.solve_(sa, ob.NewtonRaphson(ob.ConvergenceSettings(1e-10, 1e-10, 20), preconditioner=precond, cg_reltol=1e-13))
    ref = oracle.newton_solve(m, sa.load_factors(), lambda t: unit * t, oracle.ConvergenceSettings(1e-10, 1e-10, 20))
    assert sol.iterations() == ref.iterations == G.UNIAXIAL_COMPRESSION_ITERS
    for k in range(9):
        assert cases.rel_err(sol.U[k], ref.U[k]) < SOLVE_RTOL
        assert cases.rel_err(sol.F_int[k], ref.F_int[k]) < SOLVE_RTOL
    corner = int(np.argmax(mesh.xyz @ np.ones(3)))
    U = sol.U[-1].reshape(-1, 3)
    assert 1 + U[corner, 0] / 2 == pytest.approx(G.UNIAXIAL_COMPRESSION_ALPHA, rel=1e-8)
    assert 1 + U[corner, 1] == pytest.approx(G.UNIAXIAL_COMPRESSION_BETA, rel=1e-8)
    P = sol.tet_stress[-1][0].reshape(3, 3, order="F")
    assert P[0, 0] == pytest.approx(-1.0, rel=1e-4) and abs(P[1, 1]) < 1e-8 and abs(P[2, 2]) < 1e-8


def test_reference_default_linear_solver_settings(ob, oracle):
    """Un-preconditioned CG at reltol sqrt(eps) (the reference's defaults) still lands on the analytic state."""
    m, mesh = cases.box_model(4, 2, 2, mat="svk")
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (3.0, 0.0, 0.0))
    s = ob.Structure.from_arrays(m.xyz, tets=m.tets, materials=[ob.SVK(E=1.0, nu=0.3)], free_dofs=m.free_dofs, fext=lambda t: unit * t)
    sol = ob.solve(ob.NonLinearStaticAnalysis(s, NSTEPS=8), ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 30), preconditioner="none"))
    np.testing.assert_allclose(sol.U[-1], mg.homogeneous_field(mesh.xyz, 2.0, math.sqrt(0.1)), atol=1e-5)
    ref = oracle.newton_solve(m, np.linspace(1 / 8, 1, 8), lambda t: unit * t, oracle.ConvergenceSettings(1e-8, 1e-8, 30), linear="cg")
    assert sol.iterations() == ref.iterations


@pytest.mark.parametrize("name,strain_cls", [("roteng", "RotatedEngineeringStrain"), ("green", "GreenStrain")])
def test_von_mises_truss(ob, oracle, name, strain_cls):
    """examples/von_misses_truss through the object API."""
    strain = getattr(ob, strain_cls)
    m, fext, p = cases.von_mises_truss(strain.code)
    n1, n2, n3 = ob.Node(*m.xyz[0]), ob.Node(*m.xyz[1]), ob.Node(*m.xyz[2])
    d, a = math.sqrt(4 * p["A"] / math.pi), math.sqrt(p["A"])
    tl, tr = ob.Truss(n1, n2, ob.Circle(d), strain, "left_truss"), ob.Truss(n2, n3, ob.Square(a), strain, "right_truss")
    mesh = ob.Mesh(nodes=[n1, n2, n3], elements=[tl, tr])
    ob.set_dofs(mesh, "u", 3)
    bcs = ob.StructuralBoundaryCondition((ob.FixedField("u", [1, 2, 3]), [n1, n3]), (ob.FixedField("u", [2]), [n2]),
                                         (ob.GlobalLoad("u", lambda t: [0, 0, p["Fk"] * t]), [n2]))
    s = ob.Structure(mesh, ob.StructuralMaterial((ob.SVK(E=p["E"], nu=0.0, label="steel"), [tl, tr])), bcs)
    sol = ob.solve(ob.NonLinearStaticAnalysis(s, NSTEPS=5), ob.NewtonRaphson(ob.ConvergenceSettings(1e-10, 1e-10, 10), cg_reltol=1e-13, cg_maxiter=50))
    ref = oracle.newton_solve(m, np.linspace(0.2, 1, 5), fext, oracle.ConvergenceSettings(1e-10, 1e-10, 10))
    assert sol.iterations() == ref.iterations == [G.VON_MISES_ITERS[name]] * 5
    uk = sol.displacements(n2, 3)
    assert uk[-1] == pytest.approx(G.VON_MISES_UK[name], rel=1e-8)
    for k in range(5):
        assert cases.rel_err(sol.U[k], ref.U[k]) < SOLVE_RTOL
        assert cases.rel_err(sol.F_int[k], ref.F_int[k]) < SOLVE_RTOL
    assert abs(sol.displacements(n2, 1)[-1]) <= 100 * np.finfo(float).eps
    assert sol.stress(tr)[-1][0, 0] == pytest.approx(ref.truss_sig[-1][1, 0], rel=1e-8)


def test_clamped_truss_1d(ob, oracle):
    m, fext, p = cases.clamped_truss(100)
    s = ob.Structure.from_arrays(m.xyz, trusses=m.trusses, truss_area=m.truss_area, truss_strain=ob.GreenStrain,
                                 materials=[ob.SVK(E=p["E"], nu=0.3)], free_dofs=m.free_dofs, fext=fext)
    sol = ob.solve(ob.NonLinearStaticAnalysis(s, NSTEPS=10), ob.NewtonRaphson(cg_reltol=1e-13, cg_maxiter=5000))
    ref = oracle.newton_solve(m, np.linspace(0.1, 1, 10), fext, oracle.ConvergenceSettings())
    assert sol.iterations() == ref.iterations
    for k in range(10):
        assert cases.rel_err(sol.U[k], ref.U[k]) < SOLVE_RTOL
    u = sol.U[-1][-1]
    eg = 0.5 * ((p["L"] + u) ** 2 - p["L"] ** 2) / p["L"] ** 2
    assert (p["L"] + u) / p["L"] * p["E"] * eg * p["A"] == pytest.approx(p["F"], rel=1e-3)


def test_cylinder_linear_analysis_lame(ob, oracle):
    """configs[2] at oracle size: IsotropicLinearElastic cylinder as the reference ships it (LinearStaticAnalysis) and as
    a Newton analysis; Lame solution u_r = A r + B / r."""
    Ri, Re, Lz, E, nu, p = 100.0, 200.0, 30.0, 210.0, 0.3, 10.0
    mesh = mg.cylinder_tet_mesh(6, 32, 2, Ri, Re, Lz)
    fixed = {2: mesh.node_sets["z_caps"], 0: mesh.node_sets["outer_on_y_axis"], 1: mesh.node_sets["outer_on_x_axis"]}
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, fixed)
    Fp = mg.pressure_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["inner"], p)
    s = ob.Structure.from_arrays(mesh.xyz, tets=mesh.tets, materials=[ob.IsotropicLinearElastic(E, nu)], free_dofs=free, fext=lambda t: Fp * t)
    lin = ob.solve(ob.LinearStaticAnalysis(s, NSTEPS=3), ob.NewtonRaphson(cg_reltol=1e-13))
    m = oracle.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[oracle.MAT_ISOLINEAR], mat_params=[[E, nu]], free_dofs=free)
    ref = oracle.newton_solve(m, [1.0], lambda t: Fp * t, oracle.ConvergenceSettings(1e-8, 1e-8, 5))
    assert cases.rel_err(lin.U[-1], ref.U[0]) < SOLVE_RTOL
    np.testing.assert_allclose(lin.U[0] * 3, lin.U[-1], rtol=1e-7, atol=1e-12)   # linear in the load factor
    r = np.linalg.norm(mesh.xyz[:, :2], axis=1)
    ur = (lin.U[-1].reshape(-1, 3)[:, :2] * mesh.xyz[:, :2] / r[:, None]).sum(axis=1)
    A = (1 + nu) * (1 - 2 * nu) * Ri ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    B = (1 + nu) * Ri ** 2 * Re ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    np.testing.assert_allclose(ur, A * r + B / r, atol=1e-2 * (Re - Ri))
    nl = ob.solve(ob.NonLinearStaticAnalysis(s, NSTEPS=1), ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 5), cg_reltol=1e-13))
    assert nl.iterations() == ref.iterations and cases.rel_err(nl.U[-1], ref.U[0]) < SOLVE_RTOL


def test_million_tet_cube_size_independent_properties(ob):
    """configs[1] at full size (998 250 tets): properties that need no oracle -- (i) at the analytic homogeneous state the
    residual vanishes: F_int balances the lumped traction; (ii) stress / strain are the analytic P, C in every element;
    (iii) K is symmetric: x.(K y) == y.(K x); (iv) re-assembly is bitwise reproducible; (v) one Newton step from a
    perturbed state returns to the analytic solution."""
    n = 55
    mesh = mg.box_tet_mesh(n, n, n, 1.0, 1.0, 1.0)
    assert mesh.n_tets == 998_250
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    E, nu = 1.0, 0.3
    mu, K = E / 2.6, E / (3 * 0.4)
    alpha, beta = G.UNIAXIAL_COMPRESSION_ALPHA, G.UNIAXIAL_COMPRESSION_BETA   # NeoHookean, p = 1 compression
    ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[K, mu]], free_dofs=free)
    U = mg.homogeneous_field(mesh.xyz, alpha, beta)
    Fext = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (-1.0, 0.0, 0.0))
    ctx.set_U(U)
    ctx.set_Fext(Fext)
    ctx.assemble()
    Fint = ctx.get_Fint()
    mask = np.zeros(mesh.n_nodes * 3, bool)
    mask[free] = True
    assert np.abs((Fext - Fint)[mask]).max() < 1e-7 * np.abs(Fext).max()   # alpha, beta known to ~1e-11
    s, e = ctx.get_stress_strain(ob.FAMILY_TET)
    P11 = mu * alpha - mu / alpha + K * beta ** 2 * (alpha * beta ** 2 - 1)
    assert np.abs(s[:, 0] - P11).max() < 1e-9 and np.abs(s[:, 4]).max() < 1e-9 and np.abs(s[:, 1]).max() < 1e-12
    np.testing.assert_allclose(e[:, [0, 4, 8]], np.tile([alpha ** 2, beta ** 2, beta ** 2], (len(e), 1)), rtol=1e-12)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(U.size) * mask, rng.standard_normal(U.size) * mask
    Kx, Ky = ctx.spmv(x), ctx.spmv(y)
    assert abs(y @ Kx - x @ Ky) < 1e-11 * abs(y @ Kx)
    ctx.assemble()
    np.testing.assert_array_equal(ctx.get_Fint(), Fint)
    np.testing.assert_array_equal(ctx.spmv(x), Kx)
    ctx.set_U(U + 2e-5 * rng.standard_normal(U.size) * mask)   # ~1e-3 strain noise at h = 1/55
    for _ in range(4):
        info = ctx.newton_step(ob.PRECOND_JACOBI, 1e-10)
    assert np.abs(ctx.get_U() - U).max() < 1e-8 and info.norm_r / info.norm_Fext < 1e-8


def test_truss_lattice_1m_bars_size_independent_properties(ob):
    """configs[4] family at 1.0 M bars (52^3-cell braced lattice, Green strain): (i) under a homogeneous stretch every
    bar's strain / stress is the closed form of Trusses.jl:159-184 and interior nodes are in equilibrium; (ii) a rigid
    translation is in the null space of K_t at U = 0; (iii) K is symmetric; (iv) a large-displacement Newton solve of a
    tip-loaded lattice converges quadratically (residual drops below 1e-8 within the iteration budget)."""
    n = 52
    mesh = mg.truss_lattice(n, n, n, 2.0)
    assert mesh.n_bars > 1_000_000
    E, A, eps = 210e9, 2.5e-3, 1e-3
    nn = mesh.n_nodes
    ctx = ob.context_from_flat(mesh.xyz, trusses=mesh.bars, truss_area=np.full(mesh.n_bars, A), truss_strain=ob.STRAIN_GREEN,
                               mat_kind=[ob.MAT_SVK], mat_params=[[0.0, E / 2]], free_dofs=np.arange(nn * 3, dtype=np.int64))
    U = np.zeros((nn, 3))
    U[:, 0] = eps * mesh.xyz[:, 0]
    ctx.set_U(U.ravel())
    ctx.assemble()
    Fint = ctx.get_Fint().reshape(-1, 3)
    g = np.rint(mesh.xyz / 2.0).astype(int)
    interior = np.all((g > 0) & (g < n), axis=1)
    assert np.abs(Fint[interior]).max() < 1e-9 * E * A * eps
    s, e = ctx.get_stress_strain(ob.FAMILY_TRUSS)
    d = mesh.xyz[mesh.bars[:, 1]] - mesh.xyz[mesh.bars[:, 0]]
    l0 = np.linalg.norm(d, axis=1)
    l1 = np.linalg.norm(d * np.array([1 + eps, 1, 1]), axis=1)
    eg = (l1 ** 2 - l0 ** 2) / (2 * l0 ** 2)
    np.testing.assert_allclose(e[:, 0], eg, rtol=1e-10, atol=1e-16)
    np.testing.assert_allclose(s[:, 0], E * eg * l1 / l0, rtol=1e-10, atol=1e-4)
    ctx.set_U(np.zeros(nn * 3))
    ctx.assemble()
    t = np.zeros((nn, 3))
    t[:, 1] = 1.0
    assert np.abs(ctx.spmv(t.ravel())).max() < 1e-9 * E * A / 2.0
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(nn * 3), rng.standard_normal(nn * 3)
    Kx, Ky = ctx.spmv(x), ctx.spmv(y)
    assert abs(y @ Kx - x @ Ky) < 1e-11 * abs(y @ Kx)
    ctx.close()
    # clamped at x = 0, pulled at x = L: Newton on a smaller lattice of the same family (the 1 M-bar tangent needs ~1e4 CG
    # iterations per step; the property, not the size, is what this part checks)
    small = mg.truss_lattice(12, 4, 4, 2.0)
    fixed = {c: small.node_sets["x0"] for c in range(3)}
    free = mg.free_dofs_from_fixed(small.n_nodes, 3, fixed)
    F = np.zeros((small.n_nodes, 3))
    F[small.node_sets["x1"], 0] = 0.02 * E * A
    F[small.node_sets["x1"], 2] = 0.002 * E * A
    c2 = ob.context_from_flat(small.xyz, trusses=small.bars, truss_area=np.full(small.n_bars, A), truss_strain=ob.STRAIN_GREEN,
                              mat_kind=[ob.MAT_SVK], mat_params=[[0.0, E / 2]], free_dofs=free)
    c2.set_Fext(F.ravel())
    rel = []
    for _ in range(8):
        info = c2.newton_step(ob.PRECOND_JACOBI, 1e-12)
        rel.append(info.norm_r / info.norm_Fext)
    assert rel[-1] < 1e-8 and rel[-1] < rel[2] * 1e-4


def test_cylinder_600k_tets_linear_vs_lame(ob):
    """configs[2] family at 0.62 M tets ((24, 288, 15) structured cylinder, IsotropicLinearElastic): the Newton step from
    U = 0 IS the linear solve; the radial displacement matches the plane-strain Lame field to discretisation accuracy,
    reactions balance the pressure resultant, and a second Newton step changes nothing (residual at round-off)."""
    Ri, Re, Lz, E, nu, p = 100.0, 200.0, 30.0, 210.0, 0.3, 10.0
    mesh = mg.cylinder_tet_mesh(24, 288, 15, Ri, Re, Lz)
    assert mesh.n_tets == 6 * 24 * 288 * 15
    fixed = {2: mesh.node_sets["z_caps"], 0: mesh.node_sets["outer_on_y_axis"], 1: mesh.node_sets["outer_on_x_axis"]}
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, fixed)
    Fp = mg.pressure_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["inner"], p)
    ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_ISOLINEAR], mat_params=[[E, nu]], free_dofs=free)
    ctx.set_Fext(Fp)
    info = ctx.newton_step(ob.PRECOND_JACOBI, 1e-11)
    assert info.cg_residual <= info.cg_tol
    U = ctx.get_U().reshape(-1, 3)
    r = np.linalg.norm(mesh.xyz[:, :2], axis=1)
    ur = (U[:, :2] * mesh.xyz[:, :2] / r[:, None]).sum(axis=1)
    A = (1 + nu) * (1 - 2 * nu) * Ri ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    B = (1 + nu) * Ri ** 2 * Re ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    lame = A * r + B / r
    assert np.abs(ur - lame).max() < 2e-3 * lame.max()          # O(h^2) + faceted boundary at nt = 288
    assert np.abs(U[:, 2]).max() < 1e-3 * lame.max()            # plane strain: u_z = 0 up to the asymmetry of the tet split
    info2 = ctx.newton_step(ob.PRECOND_JACOBI, 1e-11)
    assert info2.norm_r / info2.norm_Fext < 1e-9                # linear material: the first step already solved it
    Fint = ctx.get_Fint().reshape(-1, 3)
    mask = np.ones(mesh.n_nodes * 3, bool)
    mask[free] = False
    # reactions: only at fixed dofs; their x / y resultants vanish (the pressure is self-equilibrated in the plane)
    Rx = (Fint - Fp.reshape(-1, 3))[:, 0][mask.reshape(-1, 3)[:, 0]].sum()
    assert abs(Rx) < 1e-6 * np.abs(Fp).sum()


def test_device_side_external_loads_match_host_apply(ob):
    """F_ext built on the device from the boundary faces / nodes (onsas_add_face_load,
    onsas_add_nodal_load, onsas_apply_loads) equals the host restatement of apply! (StructuralAnalyses.jl:228-241) --
    GlobalLoad and Pressure on TriangularFaces, GlobalLoad on nodes, duplicates summed, any load factor."""
    # (i) the shipped uniaxial example: GlobalLoad on the x = Lx faces through the object model
    s, n, t = _uniaxial_structure(ob)
    assert s.flat.load_patterns is not None and len(s.flat.load_patterns) == 3
    ctx = ob.NonLinearStaticAnalysis(s, NSTEPS=8).device_context(0)
    for lam in (0.125, 1.0, 2.5):
        s.flat.apply_loads(ctx, lam)
        np.testing.assert_allclose(ctx.get_Fext(), s.flat.fext(lam), rtol=1e-14, atol=1e-300)
    # (ii) flat API at scale: traction + pressure patterns on a 48 000-tet box and on the cylinder
    mesh = mg.box_tet_mesh(20, 20, 20, 2.0, 1.0, 1.0)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    c2 = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_SVK], mat_params=[[0.5, 0.4]], free_dofs=free)
    c2.add_face_load(mesh.faces["x1"], 0, [1.0, 0.0, 0.0])
    c2.add_face_load(mesh.faces["x1"], 0, [0.0, 0.0, 1.0])
    c2.add_face_load(mesh.faces["x1"], 1, [1.0])
    c2.add_nodal_load(mesh.node_sets["x0"][:7], [0.0, 2.0, 0.0])
    f = np.array([3.0, -0.25, 0.7, 1.5])
    c2.apply_loads(f)
    ref = (f[0] * mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (1.0, 0.0, 0.0))
           + f[1] * mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (0.0, 0.0, 1.0))
           + f[2] * mg.pressure_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], 1.0))
    nod = np.zeros((mesh.n_nodes, 3))
    nod[mesh.node_sets["x0"][:7], 1] = 2.0
    ref = ref + f[3] * nod.ravel()
    np.testing.assert_allclose(c2.get_Fext(), ref, rtol=1e-13, atol=1e-16)
    c2.apply_loads(f)                                          # deterministic: same bits again
    np.testing.assert_array_equal(c2.get_Fext(), c2.get_Fext())
    with pytest.raises(ob.OnsasError):
        c2.apply_loads(f[:2])                                  # one factor per pattern
    c2.clear_loads()
    c2.apply_loads([])
    assert not c2.get_Fext().any()
    cyl = mg.cylinder_tet_mesh(6, 32, 2)
    c3 = ob.context_from_flat(cyl.xyz, tets=cyl.tets, mat_kind=[ob.MAT_ISOLINEAR], mat_params=[[210.0, 0.3]],
                              free_dofs=np.arange(cyl.n_nodes * 3, dtype=np.int64))
    c3.add_face_load(cyl.faces["inner"], 1, [1.0])
    c3.apply_loads([10.0])
    np.testing.assert_allclose(c3.get_Fext(), mg.pressure_face_load(cyl.n_nodes, cyl.xyz, cyl.faces["inner"], 10.0), rtol=1e-13, atol=1e-13)


def test_million_tet_svk_uniaxial_extension_full_solve(ob):
    """The target sentence of BASELINE.json at size: examples/uniaxial_extension (SVK, E = 1, nu = 0.3, p = 3, Lx x Ly x Lz =
    2 x 1 x 1, NSTEPS = 8, tolerances 1e-8, max_iter 30 -- uniaxial_extension.jl:11-24,116-120) on 88 x 44 x 44 cells =
    1 022 208 tetrahedra, solved from U = 0 through Structure / NonLinearStaticAnalysis / NewtonRaphson / solve
    (NonLinearStaticAnalyses.jl:70-104).  The deformation is homogeneous, so the analytic answer of the shipped example is
    exact on every mesh: Newton iteration counts [6,5,5,4,4,4,5,5] (the reference's, = oracle direct-solve Newton on the 6-tet
    cube), alpha = 2, beta = sqrt(0.1) to 1e-8, the reactions at x = 0 sum to -p A, P and C the analytic tensors in every element."""
    import time
    mesh = mg.box_tet_mesh(88, 44, 44, 2.0, 1.0, 1.0)
    assert mesh.n_tets == 1_022_208
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (3.0, 0.0, 0.0))
    s = ob.Structure.from_arrays(mesh.xyz, tets=mesh.tets, materials=[ob.SVK(E=1.0, nu=0.3)], free_dofs=free, fext=lambda t: unit * t)
    sa = ob.NonLinearStaticAnalysis(s, NSTEPS=8)
    nr = ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 30), preconditioner="two_level", cg_reltol=1e-10)
    t0 = time.perf_counter()
    sol = ob.solve_(sa, nr)
    wall = time.perf_counter() - t0
    cg = [sum(c) for c in sol.cg_iterations]
    print(f"\n[full solve] {mesh.n_tets} tets, 8 load steps, Newton iterations {sol.iterations()}, CG iterations per step {cg}, wall {wall:.2f} s")
    assert sol.iterations() == G.UNIAXIAL_EXTENSION_ITERS == [6, 5, 5, 4, 4, 4, 5, 5]
    alpha, beta = 2.0, math.sqrt(0.1)
    U = sol.U[-1]
    Ua = mg.homogeneous_field(mesh.xyz, alpha, beta)
    assert np.abs(U - Ua).max() < 1e-8 * np.abs(Ua).max()
    corner = int(np.argmax(mesh.xyz @ np.ones(3)))
    assert 1 + U[3 * corner] / 2.0 == pytest.approx(alpha, rel=1e-8)
    assert 1 + U[3 * corner + 1] / 1.0 == pytest.approx(beta, rel=1e-8)
    Rx = sol.reactions()[-1].reshape(-1, 3)[mesh.node_sets["x0"], 0].sum()
    assert Rx == pytest.approx(-3.0, rel=1e-8)                       # -p * A, A = Ly * Lz = 1
    # P = F S with S = lambda tr(E) I + 2 G E at F = diag(alpha, beta, beta): P11 = p, P22 = P33 = 0; C = F'F
    P, Cg = sol.tet_stress[-1], sol.tet_strain[-1]
    assert np.abs(P[:, 0] - 3.0).max() < 1e-7 and np.abs(P[:, 4]).max() < 1e-7 and np.abs(P[:, 8]).max() < 1e-7
    np.testing.assert_allclose(Cg[:, [0, 4, 8]], np.tile([alpha ** 2, beta ** 2, beta ** 2], (len(Cg), 1)), rtol=1e-8)
    # every intermediate load step sits on its own analytic state too (uniaxial_extension.jl:151-154: the cubic in alpha)
    for k, lam in enumerate(sa.load_factors()):
        a_k, b_k = cases.svk_uniaxial_state(3.0 * lam, 1.0, 0.3)
        Uk = mg.homogeneous_field(mesh.xyz, a_k, b_k)
        assert np.abs(sol.U[k] - Uk).max() < 1e-7 * np.abs(Uk).max(), k


def test_gpu_solve_to_vtu(ob, tmp_path):
    """SURVEY.md 8f-2 on the GPU path: solve on the device -> Solution (flat arrays from onsas_get_stress_strain) -> write_vtk
    -> parse the .vtu: point / cell data equal what the device returned, and the cell-data label set is the reference's
    (Interfaces/VTK.jl:158-164: sigma / tau and epsilon / gamma components, 9 + 9)."""
    from onsas_jl_b200 import vtk
    m, mesh = cases.box_model(6, 3, 3, mat="neo")
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (-1.0, 0.0, 0.0))
    s = ob.Structure.from_arrays(m.xyz, tets=m.tets, materials=[ob.NeoHookean(E=1.0, nu=0.3)], free_dofs=m.free_dofs, fext=lambda t: unit * t)
    sa = ob.NonLinearStaticAnalysis(s, NSTEPS=3)
    sol = ob.solve_(sa, ob.NewtonRaphson(ob.ConvergenceSettings(1e-10, 1e-10, 20), cg_reltol=1e-13))
    ctx = sa._ctx
    sig_dev, eps_dev = ctx.get_stress_strain(ob.FAMILY_TET)      # state of the last stored step, straight from the device
    U_dev = ctx.get_U()
    path = ob.write_vtk(sol, str(tmp_path / "compression"), 3)
    arrays, meta = vtk.read_vtu(path)
    assert meta["n_points"] == mesh.n_nodes and meta["n_cells"] == mesh.n_tets
    np.testing.assert_array_equal(arrays["Points"], mesh.xyz)
    np.testing.assert_array_equal(arrays["connectivity"].reshape(-1, 4), mesh.tets)
    assert set(arrays["types"].tolist()) == {10}
    np.testing.assert_array_equal(arrays["Displacement"], U_dev.reshape(-1, 3))
    sig = sig_dev.reshape(-1, 3, 3).transpose(0, 2, 1)           # column-major 3x3 -> [e, i, j]
    eps = eps_dev.reshape(-1, 3, 3).transpose(0, 2, 1)
    ij = {"x": 0, "y": 1, "z": 2}
    for lab in vtk.STRESS_LABELS:
        np.testing.assert_array_equal(arrays[lab], sig[:, ij[lab[1]], ij[lab[2]]])
    for lab in vtk.STRAIN_LABELS:
        np.testing.assert_array_equal(arrays[lab], eps[:, ij[lab[1]], ij[lab[2]]])
    assert {k for k in arrays if k[0] in "στϵγ"} == set(vtk.STRESS_LABELS) | set(vtk.STRAIN_LABELS)
    assert vtk.STRESS_LABELS == ["σxx", "σyy", "σzz", "τyz", "τxz", "τxy", "τzy", "τzx", "τyx"]      # VTK.jl:158-161
    assert vtk.STRAIN_LABELS == ["ϵxx", "ϵyy", "ϵzz", "γyz", "γxz", "γxy", "γzy", "γzx", "γyx"]      # VTK.jl:162-164
    # analytic check of what landed on disk: P11 = -1 (the traction), P22 = 0, homogeneous in every cell
    assert np.abs(arrays["σxx"] + 1.0).max() < 1e-8 and np.abs(arrays["σyy"]).max() < 1e-8
    pvd = ob.write_vtk(sol, str(tmp_path / "series"))
    assert open(pvd, encoding="utf-8").read().count("<DataSet") == 3
