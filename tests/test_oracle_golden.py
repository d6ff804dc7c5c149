"""Pins the CPU oracle against every literal vector / closed form the reference's tests hold for the hot path
(SURVEY.md section 4 table).  Tolerances are the reference's own (rtol 1e-3 on 4-5 printed digits) unless exact."""
import numpy as np
import pytest

from tests.golden import reference_vectors as G

RTOL = 1e-3  # test/entities/tetrahedrons.jl:9, test/materials/materials.jl


def test_tet_volume_exact(oracle):
    assert oracle.tet_volume(G.TET_NODES) == G.TET_VOLUME  # test/entities/tetrahedrons.jl:66 (==)


def test_tet_svk_golden(oracle):
    f, K, P, Cc = oracle.tet_internal_forces(oracle.MAT_SVK, G.TET_LAMBDA, G.TET_G, G.TET_NODES, G.TET_U)
    np.testing.assert_allclose(f, G.TET_F_INT, rtol=RTOL)       # :97
    np.testing.assert_allclose(K, G.TET_K, rtol=RTOL)           # :98
    np.testing.assert_allclose(Cc, G.TET_C, rtol=RTOL)          # :99
    # norm-wise like Julia's isapprox, and symmetric
    assert np.linalg.norm(K - G.TET_K) <= RTOL * np.linalg.norm(G.TET_K)
    np.testing.assert_array_equal(K, K.T)


def test_tet_isolinear_equals_svk_at_zero(oracle):
    """test/entities/tetrahedrons.jl:122-138."""
    Emod = G.TET_G * (3 * G.TET_LAMBDA + 2 * G.TET_G) / (G.TET_LAMBDA + G.TET_G)
    # the reference builds IsotropicLinearElastic(E, shear_modulus) -- it passes G where nu is expected (:124-125);
    # reproduce that call literally
    nu = G.TET_G
    f, K, s, e = oracle.tet_internal_forces(oracle.MAT_ISOLINEAR, Emod, nu, G.TET_NODES, G.TET_U)
    lam = Emod * nu / ((1 + nu) * (1 - 2 * nu))
    Gs = Emod / (2 * (1 + nu))
    _, Ks, _, _ = oracle.tet_internal_forces(oracle.MAT_SVK, lam, Gs, G.TET_NODES, np.zeros(12))
    np.testing.assert_allclose(Ks @ G.TET_U, f, rtol=RTOL)
    np.testing.assert_allclose(Ks, K, rtol=RTOL, atol=1e-12)


def test_negative_volume_is_an_error(oracle):
    X = G.TET_NODES[[1, 0, 2, 3]]
    with pytest.raises(oracle.NegativeVolumeError):
        oracle.tet_internal_forces(oracle.MAT_SVK, 1.0, 1.0, X, np.zeros(12))


def test_svk_stress_golden(oracle):
    S, D = oracle.material_stress(oracle.MAT_SVK, G.MAT_LAMBDA, G.MAT_G, G.MAT_E)
    np.testing.assert_allclose(S, G.SVK_S, rtol=RTOL)   # test/materials/materials.jl:128-129
    np.testing.assert_allclose(D, G.SVK_D, rtol=RTOL, atol=1e-15)


def test_isolinear_stress_closed_form(oracle):
    """test/materials/materials.jl:43-74: sigma = lambda tr(eps) I + 2 G eps."""
    E, nu = 210e9, 0.3
    eps = np.array([[1e-4, 2e-5, 0], [2e-5, -3e-5, 1e-5], [0, 1e-5, 5e-5]])
    s, D = oracle.material_stress(oracle.MAT_ISOLINEAR, E, nu, eps)
    lam, Gs = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    np.testing.assert_allclose(s, lam * np.trace(eps) * np.eye(3) + 2 * Gs * eps, rtol=1e-14)
    assert D[0, 0] == lam + 2 * Gs and D[0, 1] == lam and D[3, 3] == Gs


def _num_grad_S(oracle, kind, p0, p1, E, h=1e-6):
    """central differences of S wrt symmetric perturbations, in the reference's Voigt layout."""
    VI, VJ = [0, 1, 2, 1, 0, 0], [0, 1, 2, 2, 2, 1]
    D = np.zeros((6, 6))
    for b in range(6):
        dE = np.zeros((3, 3))
        dE[VI[b], VJ[b]] += 1
        dE[VJ[b], VI[b]] += 1
        if VI[b] == VJ[b]:
            dE /= 2
        Sp, _ = oracle.material_stress(kind, p0, p1, E + h * dE)
        Sm, _ = oracle.material_stress(kind, p0, p1, E - h * dE)
        dS = (Sp - Sm) / (2 * h)
        # perturbing E_kl and E_lk together changes S by 2*DD_ijkl*h for k != l
        scale = 1.0 if VI[b] == VJ[b] else 0.5
        for a in range(6):
            D[a, b] = dS[VI[a], VJ[a]] * scale
    return D


@pytest.mark.parametrize("kind,p0,p1", [(0, G.MAT_LAMBDA, G.MAT_G), (1, G.MAT_K, G.MAT_G)])
def test_tangent_is_derivative_of_stress(oracle, kind, p0, p1):
    """NeoHookean: the reference gets dS/dE from ForwardDiff (NeoHookeanMaterial.jl:115-129) and pins it only by
    self-consistency with AD of the strain energy (test/materials/materials.jl:141-186, rtol 1e-3)."""
    S, D = oracle.material_stress(kind, p0, p1, G.MAT_E)
    Dn = _num_grad_S(oracle, kind, p0, p1, G.MAT_E)
    np.testing.assert_allclose(D, Dn, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(D, D.T, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("kind,p0,p1", [(0, G.MAT_LAMBDA, G.MAT_G), (1, G.MAT_K, G.MAT_G)])
def test_stress_is_derivative_of_energy(oracle, kind, p0, p1):
    """S = dPsi/dE: the consistency the reference checks through its generic HyperElastic material
    (test/materials/materials.jl:123-186)."""
    S, _ = oracle.material_stress(kind, p0, p1, G.MAT_E)
    h = 1e-6
    Sn = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            dE = np.zeros((3, 3))
            dE[i, j] += 0.5
            dE[j, i] += 0.5
            Sn[i, j] = (oracle.strain_energy(kind, p0, p1, G.MAT_E + h * dE) -
                        oracle.strain_energy(kind, p0, p1, G.MAT_E - h * dE)) / (2 * h)
    np.testing.assert_allclose(S, Sn, rtol=1e-7, atol=1e-9)


def test_tet_tangent_is_derivative_of_force(oracle):
    """K_e = d f_int / d u_e for both hyperelastic materials (consistency of the whole element restatement)."""
    for kind, p0, p1 in [(0, G.TET_LAMBDA, G.TET_G), (1, G.MAT_K, G.MAT_G)]:
        u = G.TET_U * 0.2
        _, K, _, _ = oracle.tet_internal_forces(kind, p0, p1, G.TET_NODES, u)
        h = 1e-6
        Kn = np.zeros((12, 12))
        for c in range(12):
            du = np.zeros(12)
            du[c] = h
            fp = oracle.tet_internal_forces(kind, p0, p1, G.TET_NODES, u + du)[0]
            fm = oracle.tet_internal_forces(kind, p0, p1, G.TET_NODES, u - du)[0]
            Kn[:, c] = (fp - fm) / (2 * h)
        np.testing.assert_allclose(K, Kn, rtol=1e-6, atol=1e-8)


def test_truss_1d_rotated_engineering(oracle):
    """test/entities/trusses.jl:17-61."""
    E, nu, A = 1.0, 0.3, 1.0
    X = np.array([[-1.0], [1.0]])
    u = np.array([0.1, 0.25])
    l_ref, l_def = 2.0, (1 + 0.25) - (-1 + 0.1)
    f, K, s, e = oracle.truss_internal_forces(oracle.STRAIN_ROTENG, 1, E, A, X, u)
    eps = (l_def ** 2 - l_ref ** 2) / (l_ref * (l_ref + l_def))
    assert e[0, 0] == pytest.approx(eps, rel=RTOL)
    assert s[0, 0] == pytest.approx(E * eps * l_def / l_ref, rel=RTOL)
    assert f[0] == pytest.approx(-E * eps * A, rel=RTOL) and f[1] == pytest.approx(E * eps * A, rel=RTOL)
    np.testing.assert_allclose(K, K[0, 0] * np.array([[1, -1], [-1, 1]]), rtol=RTOL)


def test_truss_3d_at_rest(oracle):
    """test/entities/trusses.jl:63-124: f = 0, K[1] == E A / l, zero stress / strain at u = 0."""
    X = np.array([[-1.0, 0, 0], [1.0, 0, 0]])
    f, K, s, e = oracle.truss_internal_forces(oracle.STRAIN_ROTENG, 3, 1.0, 1.0, X, np.zeros(6))
    assert np.linalg.norm(f) == 0 and K[0, 0] == 1.0 * 1.0 / 2.0 and s[0, 0] == 0 and e[0, 0] == 0


@pytest.mark.parametrize("strain", [0, 1])
def test_truss_tangent_is_derivative_of_force(oracle, strain):
    rng = np.random.default_rng(3)
    X = np.array([[0.0, 0.1, -0.2], [1.0, 0.7, 0.4]])
    u = rng.uniform(-0.2, 0.2, 6)
    f, K, _, _ = oracle.truss_internal_forces(strain, 3, 2.0, 0.5, X, u)
    h = 1e-6
    Kn = np.zeros((6, 6))
    for c in range(6):
        du = np.zeros(6)
        du[c] = h
        Kn[:, c] = (oracle.truss_internal_forces(strain, 3, 2.0, 0.5, X, u + du)[0] -
                    oracle.truss_internal_forces(strain, 3, 2.0, 0.5, X, u - du)[0]) / (2 * h)
    np.testing.assert_allclose(K, Kn, rtol=1e-6, atol=1e-8)


def test_assembler_coo_order_and_sum(oracle):
    """test/structural_solvers/structural_solvers.jl:77-126: two 2-dof elements with the same K_e.
    Restated with 1-D trusses: EA/l = 1 at rest gives K_e = [1 -1; -1 1]; the reference's literal
    I/J order is checked on the triplet generator, the sum on the assembled matrix."""
    Ke = G.COO_KE
    I, J, V = [], [], []
    for dofs in G.COO_DOFS:  # Assemblers.jl:52-67
        for c in range(2):
            for r in range(2):
                I.append(dofs[r])
                J.append(dofs[c])
                V.append(Ke[r, c])
    assert I == G.COO_I and J == G.COO_J
    Kg = np.zeros((3, 3))
    for i, j, v in zip(I, J, V):
        Kg[i - 1, j - 1] += v
    np.testing.assert_array_equal(Kg, G.COO_KGLOB)
    # the C assembly on a 2-bar chain: middle dof gets the sum of both elements
    m = oracle.FlatModel(xyz=[[0.0], [1.0], [2.0]], dim=1, trusses=[[0, 1], [1, 2]], truss_area=[1.0, 1.0],
                         mat_kind=[0], mat_params=[[0.0, 0.5]], free_dofs=[1, 2])
    asm = oracle.Assembly(m).assemble(np.zeros(3))
    np.testing.assert_allclose(asm.dense(), np.array([[1, -1, 0], [-1, 2, -1], [0, -1, 1.0]]), rtol=1e-15)


def test_two_truss_state_assembly(oracle):
    """test/structural_analyses/static_analyses.jl:150-188: F_int (9) and K (9x9) of two trusses vs the dense manual sum."""
    from tests.cases import von_mises_truss
    m, _, _ = von_mises_truss(oracle.STRAIN_ROTENG)
    U = np.array([0, 0, 0, 0.01, 0, -0.02, 0, 0, 0.0])
    asm = oracle.Assembly(m).assemble(U)
    f, K, _, _ = oracle.eval_trusses(m, U)
    Kd = np.zeros((9, 9))
    Fd = np.zeros(9)
    for e, (a, b) in enumerate(m.trusses):
        dofs = [3 * a, 3 * a + 1, 3 * a + 2, 3 * b, 3 * b + 1, 3 * b + 2]
        Kd[np.ix_(dofs, dofs)] += K[e].reshape(6, 6, order="F")
        Fd[dofs] += f[e]
    np.testing.assert_allclose(asm.dense(), Kd, rtol=1e-14, atol=1e-3)
    np.testing.assert_allclose(asm.F_int, Fd, rtol=1e-14, atol=1e-6)


def test_cg_restatement_matches_direct_solve(oracle):
    from tests.cases import box_model, random_U
    m, _ = box_model(4, 2, 2, jitter=0.1)
    asm = oracle.Assembly(m).assemble(random_U(m, 0.02))
    mask = m.free_mask()
    b = np.random.default_rng(5).standard_normal(m.n_dofs) * mask
    import scipy.sparse.linalg as spla
    A = asm.csr()[m.free_dofs][:, m.free_dofs].tocsc()
    xd = spla.spsolve(A, b[m.free_dofs])
    for diag in (None, np.where(mask, asm.csr().diagonal(), 1.0)):
        x, it, res = oracle.cg(asm.rowptr, asm.col, asm.val, mask, b, diag=diag, reltol=1e-12)
        np.testing.assert_allclose(x[m.free_dofs], xd, rtol=1e-8, atol=1e-10)
        assert 0 < it <= len(m.free_dofs)
    # default tolerance sqrt(eps) and zero rhs -> zero iterations (tolerance = 0, residual 0 <= 0)
    x, it, res = oracle.cg(asm.rowptr, asm.col, asm.val, mask, np.zeros(m.n_dofs))
    assert it == 0 and res == 0.0


def test_mt_assembly_matches_serial(oracle):
    from tests.cases import box_model, random_U
    m, _ = box_model(5, 3, 2, mat="neo", jitter=0.1)
    U = random_U(m)
    a1 = oracle.Assembly(m).assemble(U)
    a2 = oracle.AssemblyMT(m).assemble(U)
    np.testing.assert_array_equal(a1.val, a2.val)  # same per-entry summation order -> bitwise
    np.testing.assert_array_equal(a1.F_int, a2.F_int)
