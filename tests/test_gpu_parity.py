"""Parity tests proper: the CUDA path (through the C ABI) against the oracle on the same seeded inputs.
Bars (north star): per-element f_int / K_t within 1e-10 relative; converged displacements, reactions and Newton
iteration counts within 1e-8 relative of the direct-solve result.  Run with `pytest -m gpu` on a B200."""
import math

import numpy as np
import pytest

from onsas_jl_b200 import meshgen as mg
from tests import cases
from tests.golden import reference_vectors as G

pytestmark = pytest.mark.gpu

ELEM_RTOL = 1e-10   # north star: per-element f_int and K_t
SOLVE_RTOL = 1e-8   # north star: converged displacements / reactions


def _ctx(ob, m, **kw):
    return ob.context_from_flat(m.xyz, tets=m.tets, trusses=m.trusses, truss_area=m.truss_area,
                                truss_strain=m.truss_strain, mat_kind=m.mat_kind, mat_params=m.mat_params,
                                tet_mat=m.tet_mat, truss_mat=m.truss_mat, free_dofs=m.free_dofs, **kw)


def _rowwise_rel(a, b):
    scale = np.abs(b).max(axis=1, keepdims=True)
    return float((np.abs(a - b) / scale).max())


def test_golden_tet_through_the_abi(ob):
    """test/entities/tetrahedrons.jl:73-99 literal vectors, rtol 1e-3 as in the reference."""
    ctx = ob.context_from_flat(G.TET_NODES, tets=[[0, 1, 2, 3]], mat_kind=[ob.MAT_SVK], mat_params=[[G.TET_LAMBDA, G.TET_G]])
    ctx.set_U(G.TET_U)
    f, K, s, e = ctx.eval_elements(ob.FAMILY_TET)
    np.testing.assert_allclose(f[0], G.TET_F_INT, rtol=1e-3)
    np.testing.assert_allclose(K[0].reshape(12, 12, order="F"), G.TET_K, rtol=1e-3)
    np.testing.assert_allclose(e[0].reshape(3, 3, order="F"), G.TET_C, rtol=1e-3)
    # assembled path on the same single element
    ctx.assemble()
    np.testing.assert_allclose(ctx.get_Fint(), G.TET_F_INT, rtol=1e-3)
    rp, ci, v = ctx.get_csr()
    import scipy.sparse as sp
    np.testing.assert_allclose(sp.csr_matrix((v, ci, rp), shape=(12, 12)).toarray(), G.TET_K, rtol=1e-3)


@pytest.mark.parametrize("mat", ["svk", "neo", "iso"])
def test_per_element_parity_1e5_random_tets(ob, oracle, mat):
    """SURVEY.md 8d per-element parity set: 1e5 random distorted tets, f_e / K_e / stress / strain within 1e-10."""
    m, U = cases.random_tet_model(100_000, mat)
    f, K, s, e = oracle.eval_tets(m, U)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    f2, K2, s2, e2 = ctx.eval_elements(ob.FAMILY_TET)
    # The 1e-10 bar is for states where the reference arithmetic itself is well conditioned.  The random set contains
    # nearly collapsed / inverted elements (J = |det F| down to 3e-4, 1 % have J < 0.1); for NeoHookean the reference
    # inverts C = 2E + I there (cond(C) ~ 1/J^2), so its own result carries ~eps/J^2 error.  Elements with J >= 0.1
    # must meet 1e-10; the rest are held to the conditioning-scaled bound 1e-10 * (0.1/J)^2.
    J = np.sqrt(np.abs(np.linalg.det(e.reshape(-1, 3, 3)))) if mat != "iso" else np.ones(len(e))
    tol = (ELEM_RTOL * np.maximum(1.0, (0.1 / J) ** 2))[:, None] if mat == "neo" else ELEM_RTOL

    def ok(a, b):
        return bool(np.all(np.abs(a - b) / np.abs(b).max(axis=1, keepdims=True) < tol))
    assert ok(f2, f) and ok(K2, K) and ok(s2, s) and ok(e2, e)
    well = J >= 0.1
    assert well.mean() > 0.98
    assert _rowwise_rel(f2[well], f[well]) < ELEM_RTOL and _rowwise_rel(K2[well], K[well]) < ELEM_RTOL
    # chunked access returns the same numbers
    f3, K3, _, _ = ctx.eval_elements(ob.FAMILY_TET, first=777, count=1000)
    np.testing.assert_array_equal(f3, f2[777:1777])
    np.testing.assert_array_equal(K3, K2[777:1777])
    # the fused assembly kernel evaluates the same rows: disconnected tets -> K is block diagonal = K_e
    ctx.assemble()
    assert ok(ctx.get_Fint().reshape(-1, 12), f)
    s4, e4 = ctx.get_stress_strain(ob.FAMILY_TET)
    assert ok(s4, s) and ok(e4, e)
    rp, ci, v = ctx.get_csr()
    assert np.all(np.diff(rp) == 12)
    Kasm = v.reshape(-1, 12, 12)            # 12 rows of 12 entries per tet, row-major
    Kref = K.reshape(-1, 12, 12).transpose(0, 2, 1)  # oracle is column-major
    assert ok(Kasm.reshape(-1, 144), Kref.reshape(-1, 144))
    # the fused kernel and the per-element kernel run the same device arithmetic
    assert _rowwise_rel(Kasm.reshape(-1, 144), K2.reshape(-1, 12, 12).transpose(0, 2, 1).reshape(-1, 144)) < 1e-13


@pytest.mark.parametrize("strain", [0, 1])
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_per_element_parity_trusses(ob, oracle, strain, dim):
    rng = np.random.default_rng(11)
    n = 20_000
    xyz = rng.uniform(0, 1, (2 * n, dim))
    xyz[1::2] += 0.5
    bars = np.arange(2 * n, dtype=np.int32).reshape(n, 2)
    m = oracle.FlatModel(xyz=xyz, dim=dim, trusses=bars, truss_area=rng.uniform(0.5, 2, n), truss_strain=strain,
                         mat_kind=[0, 1], mat_params=[[0.4, 1.1], [2.0, 0.7]], truss_mat=rng.integers(0, 2, n),
                         free_dofs=np.arange(2 * n * dim))
    U = rng.uniform(-0.2, 0.2, 2 * n * dim)
    f, K, s, e = oracle.eval_trusses(m, U)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    f2, K2, s2, e2 = ctx.eval_elements(ob.FAMILY_TRUSS)
    assert _rowwise_rel(f2, f) < ELEM_RTOL and _rowwise_rel(K2, K) < ELEM_RTOL
    assert cases.rel_err(s2, s) < ELEM_RTOL and cases.rel_err(e2, e) < ELEM_RTOL
    ctx.assemble()
    assert cases.rel_err(ctx.get_Fint(), f.ravel()) < ELEM_RTOL
    s3, e3 = ctx.get_stress_strain(ob.FAMILY_TRUSS)
    assert cases.rel_err(s3, s) < ELEM_RTOL and cases.rel_err(e3, e) < ELEM_RTOL


@pytest.mark.parametrize("mat,grid", [("svk", (12, 6, 6)), ("neo", (7, 5, 3)), ("iso", (9, 2, 1)), ("svk", (1, 1, 1))])
def test_assembled_system_matches_reference_order_assembly(ob, oracle, mat, grid):
    m, _ = cases.box_model(*grid, mat=mat, jitter=0.15)
    U = cases.random_U(m)
    ref = oracle.Assembly(m).assemble(U)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    ctx.assemble()
    rp, ci, v = ctx.get_csr()
    np.testing.assert_array_equal(rp, ref.rowptr)
    np.testing.assert_array_equal(ci, ref.col)
    # SURVEY.md 8c summation-order caveat: compare against the row's absolute scale
    row_scale = np.repeat(np.maximum.reduceat(np.abs(ref.val), ref.rowptr[:-1]), np.diff(ref.rowptr))
    assert float((np.abs(v - ref.val) / row_scale).max()) < 1e-12
    assert cases.rel_err(ctx.get_Fint(), ref.F_int) < 1e-12
    s, e = ctx.get_stress_strain(ob.FAMILY_TET)
    assert _rowwise_rel(s, ref.tet_sig) < ELEM_RTOL and _rowwise_rel(e, ref.tet_eps) < ELEM_RTOL


def test_assembly_is_bitwise_deterministic_and_launch_independent(ob, oracle):
    m, _ = cases.box_model(16, 8, 8, mat="neo", jitter=0.1)
    U = cases.random_U(m)
    runs = []
    for minb in (2, 2, 1, 3):
        ctx = _ctx(ob, m)
        ctx.set_option(ob._lib.OPT_ASM_MINBLOCKS, minb)
        ctx.set_U(U)
        ctx.assemble()
        ctx.assemble()   # re-assembly overwrites, it does not accumulate (reset_assemble!, StaticAnalyses.jl:125-132)
        runs.append((ctx.get_csr()[2], ctx.get_Fint()))
        ctx.close()
    for v, F in runs[1:]:
        np.testing.assert_array_equal(v, runs[0][0])
        np.testing.assert_array_equal(F, runs[0][1])


def test_mixed_materials_and_families(ob, oracle):
    m, mesh = cases.box_model(5, 3, 3, jitter=0.1)
    rng = np.random.default_rng(2)
    bars = np.stack([np.arange(0, mesh.n_nodes - 1), np.arange(1, mesh.n_nodes)], axis=1).astype(np.int32)
    mm = oracle.FlatModel(xyz=m.xyz, tets=m.tets, tet_mat=rng.integers(0, 3, len(m.tets)), trusses=bars,
                          truss_mat=rng.integers(0, 2, len(bars)), truss_area=rng.uniform(0.1, 0.3, len(bars)),
                          truss_strain=1, mat_kind=[0, 1, 2], mat_params=[[0.58, 0.38], [0.83, 0.38], [1.0, 0.3]],
                          free_dofs=m.free_dofs)
    U = cases.random_U(mm, 0.03)
    ref = oracle.Assembly(mm).assemble(U)
    ctx = _ctx(ob, mm)
    ctx.set_U(U)
    ctx.assemble()
    rp, ci, v = ctx.get_csr()
    np.testing.assert_array_equal(ci, ref.col)
    assert cases.rel_err(v, ref.val) < 1e-12 and cases.rel_err(ctx.get_Fint(), ref.F_int) < 1e-12
    s, e = ctx.get_stress_strain(ob.FAMILY_TRUSS)
    assert cases.rel_err(s, ref.truss_sig) < 1e-12 and cases.rel_err(e, ref.truss_eps) < 1e-12


def test_negative_volume_is_reported(ob):
    X = G.TET_NODES[[1, 0, 2, 3]]
    with pytest.raises(ob.NegativeVolumeError, match="Element with negative volume, check connectivity."):
        ob.context_from_flat(X, tets=[[0, 1, 2, 3]], mat_kind=[0], mat_params=[[1.0, 1.0]])


def test_abi_argument_errors(ob):
    ctx = ob.DeviceContext(0)
    with pytest.raises(ob.OnsasError) as ei:
        ctx.assemble()
    assert ei.value.status == ob._lib.ERR_NOT_READY
    ctx.set_nodes(np.zeros((4, 3)))
    ctx.set_materials([0], [[1.0, 1.0]])
    ctx.set_tets([[0, 1, 2, 9]])
    ctx.set_free_dofs(np.arange(12))
    with pytest.raises(ob.OnsasError) as ei:
        ctx.finalize()
    assert ei.value.status == ob._lib.ERR_INVALID_ARG
    with pytest.raises(ob.OnsasError):
        ctx.set_materials([7], [[1.0, 1.0]])


# ----------------------------------------------------------------------------- linear solve

@pytest.mark.parametrize("cg_mode", [0, 1, 2])
@pytest.mark.parametrize("precond", [0, 1])
def test_spmv_and_pcg_match_oracle(ob, oracle, cg_mode, precond):
    m, _ = cases.box_model(10, 5, 5, jitter=0.1)
    U = cases.random_U(m, 0.02)
    ref = oracle.Assembly(m).assemble(U)
    ctx = _ctx(ob, m)
    ctx.set_option(ob._lib.OPT_CG_MODE, cg_mode)
    ctx.set_U(U)
    ctx.assemble()
    mask = m.free_mask()
    rng = np.random.default_rng(4)
    x = rng.standard_normal(m.n_dofs) * mask
    assert cases.rel_err(ctx.spmv(x), (ref.csr() @ x) * mask) < 1e-13
    b = rng.standard_normal(m.n_dofs)
    diag = np.where(mask, ref.csr().diagonal(), 1.0) if precond else None
    import scipy.sparse.linalg as spla
    A = ref.csr()
    xd = np.zeros(m.n_dofs)
    xd[m.free_dofs] = spla.spsolve(A[m.free_dofs][:, m.free_dofs].tocsc(), b[m.free_dofs])
    for reltol in (None, 1e-12):
        xo, ito, reso = oracle.cg(ref.rowptr, ref.col, ref.val, mask, b, diag=diag, reltol=reltol)
        xs, its, res = ctx.pcg(b, precond, reltol)
        assert abs(its - ito) <= max(2, ito // 50)
        tol = (reltol or math.sqrt(np.finfo(float).eps))
        nb = np.linalg.norm(b * mask)
        assert res <= tol * nb * (1 + 1e-12)
        assert np.linalg.norm((A @ xs - b) * mask) <= 10 * tol * nb          # true residual
        assert cases.rel_err(xs, xo) < 1e4 * tol and np.all(xs[mask == 0] == 0)
    assert cases.rel_err(xs, xd) < 1e-8                                        # tight solve == direct solve
    # deterministic: the same solve twice gives the same bits
    x1, i1, _ = ctx.pcg(b, precond, 1e-12)
    x2, i2, _ = ctx.pcg(b, precond, 1e-12)
    assert i1 == i2
    np.testing.assert_array_equal(x1, x2)
    # maxiter is honoured, zero rhs takes zero iterations (IterativeSolvers: residual 0 <= tol 0)
    assert ctx.pcg(b, precond, 1e-12, maxiter=3)[1] == 3
    assert ctx.pcg(np.zeros(m.n_dofs), precond)[1] == 0


def test_persistent_and_multilaunch_cg_agree(ob, oracle):
    m, _ = cases.box_model(8, 4, 4, mat="neo", jitter=0.1)
    U = cases.random_U(m, 0.02)
    b = np.random.default_rng(1).standard_normal(m.n_dofs)
    out = []
    for mode in (0, 1, 2):
        ctx = _ctx(ob, m)
        ctx.set_option(ob._lib.OPT_CG_MODE, mode)
        ctx.set_U(U)
        ctx.assemble()
        out.append(ctx.pcg(b, ob.PRECOND_JACOBI, 1e-12))
    for o in out[1:]:
        assert abs(out[0][1] - o[1]) <= 1 and cases.rel_err(out[0][0], o[0]) < 1e-9




def test_high_valence_hub_falls_back_to_the_register_fed_solver(ob, oracle):
    """A hub node with 400 bars: its slice has more (row, element) pairs than the assembly CTA has threads (the pair
    loop wraps) and its row is 401 blocks wide -- too wide for the shared-memory ring of the streamed CG, which must
    hand over to the register-fed persistent kernel.  Assembly, SpMV and all three solver modes still match the
    oracle."""
    rng = np.random.default_rng(5)
    n_spokes = 400
    xyz = np.vstack([np.zeros((1, 3)), rng.standard_normal((n_spokes, 3)) + np.array([0.0, 0.0, 3.0])])
    xyz[1:] /= np.linalg.norm(xyz[1:], axis=1, keepdims=True) / rng.uniform(1.0, 2.0, (n_spokes, 1))
    bars = np.stack([np.zeros(n_spokes, np.int32), np.arange(1, n_spokes + 1, dtype=np.int32)], axis=1)
    ring = np.stack([np.arange(1, n_spokes, dtype=np.int32), np.arange(2, n_spokes + 1, dtype=np.int32)], axis=1)
    bars = np.vstack([bars, ring]).astype(np.int32)
    free = np.arange(3, dtype=np.int64)                      # only the hub moves
    m = oracle.FlatModel(xyz=xyz, trusses=bars, truss_area=rng.uniform(0.5, 1.5, len(bars)), truss_strain=1,
                         mat_kind=[0], mat_params=[[0.0, 50.0]], free_dofs=free)
    U = np.zeros(m.n_dofs)
    U[:3] = [0.05, -0.02, 0.08]
    ref = oracle.Assembly(m).assemble(U)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    ctx.assemble()
    rp, ci, v = ctx.get_csr()
    np.testing.assert_array_equal(ci, ref.col)
    assert cases.rel_err(v, ref.val) < 1e-12 and cases.rel_err(ctx.get_Fint(), ref.F_int) < 1e-12
    mask = m.free_mask()
    x = rng.standard_normal(m.n_dofs) * mask
    assert cases.rel_err(ctx.spmv(x), (ref.csr() @ x) * mask) < 1e-13
    b = rng.standard_normal(m.n_dofs)
    A = ref.csr()
    xd = np.zeros(m.n_dofs)
    xd[free] = np.linalg.solve(A[free][:, free].toarray(), b[free])
    for mode in (0, 1, 2):
        ctx.set_option(ob._lib.OPT_CG_MODE, mode)
        xs, its, res = ctx.pcg(b, ob.PRECOND_JACOBI, 1e-13)
        assert its <= 4 and cases.rel_err(xs, xd) < 1e-10


@pytest.mark.parametrize("mat", ["svk", "neo"])
def test_random_node_and_element_numbering(ob, oracle, mat):
    """An 'unstructured' numbering of the same mesh: nodes and elements randomly permuted, so rows have scattered
    columns, slices have ragged widths and a node's elements are far apart in memory.  Assembly (reference order =
    ascending element id of the PERMUTED mesh), SpMV and the three solver modes still match the oracle, and one
    Newton step equals the direct solve."""
    m0, _ = cases.box_model(11, 6, 5, mat=mat, jitter=0.15)
    rng = np.random.default_rng(11)
    n = m0.xyz.shape[0]
    perm = rng.permutation(n)                      # new id of old node i
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    xyz = m0.xyz[inv]
    tets = perm[m0.tets][rng.permutation(len(m0.tets))].astype(np.int32)
    free = np.sort(perm[m0.free_dofs // 3] * 3 + m0.free_dofs % 3)
    m = oracle.FlatModel(xyz=xyz, tets=tets, mat_kind=m0.mat_kind, mat_params=m0.mat_params, free_dofs=free)
    U = cases.random_U(m, 0.03)
    ref = oracle.Assembly(m).assemble(U)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    ctx.assemble()
    rp, ci, v = ctx.get_csr()
    np.testing.assert_array_equal(ci, ref.col)
    assert cases.rel_err(v, ref.val) < 1e-12 and cases.rel_err(ctx.get_Fint(), ref.F_int) < 1e-12
    mask = m.free_mask()
    x = rng.standard_normal(m.n_dofs) * mask
    assert cases.rel_err(ctx.spmv(x), (ref.csr() @ x) * mask) < 1e-13
    b = rng.standard_normal(m.n_dofs)
    import scipy.sparse.linalg as spla
    A = ref.csr()
    xd = np.zeros(m.n_dofs)
    xd[free] = spla.spsolve(A[free][:, free].tocsc(), b[free])
    its = []
    for mode in (0, 1, 2):
        ctx.set_option(ob._lib.OPT_CG_MODE, mode)
        xs, it, res = ctx.pcg(b, ob.PRECOND_JACOBI, 1e-12)
        assert cases.rel_err(xs, xd) < 1e-8
        its.append(it)
    assert max(its) - min(its) <= max(2, min(its) // 50)
    ctx.set_option(ob._lib.OPT_CG_MODE, 0)
    Fext = rng.standard_normal(m.n_dofs) * 1e-3
    ctx.set_Fext(Fext)
    info = ctx.newton_step(ob.PRECOND_JACOBI, 1e-13)
    Uref = U.copy()
    Uref[free] += spla.spsolve(A[free][:, free].tocsc(), (Fext - ref.F_int)[free])
    assert cases.rel_err(ctx.get_U(), Uref) < 1e-8


def test_two_level_preconditioner(ob, oracle):
    """Jacobi + aggregated coarse space (precond = 2).  Same solution as the direct solve, far fewer
    iterations than Jacobi on a mesh with several aggregates, bitwise reproducible, coarse inverse refreshed after a
    re-assembly, and rejected outside the streamed solver."""
    m, _ = cases.box_model(24, 12, 12, mat="neo", jitter=0.1)      # 4225 nodes -> 12 aggregates
    U = cases.random_U(m, 0.002)
    ref = oracle.Assembly(m).assemble(U)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    ctx.assemble()
    rng = np.random.default_rng(8)
    b = rng.standard_normal(m.n_dofs)
    import scipy.sparse.linalg as spla
    A = ref.csr()
    free = m.free_dofs
    xd = np.zeros(m.n_dofs)
    xd[free] = spla.spsolve(A[free][:, free].tocsc(), b[free])
    xj, itj, _ = ctx.pcg(b, ob.PRECOND_JACOBI, 1e-12)
    # translations only, then (default) translations + rotations of every aggregate
    ctx.set_option(ob._lib.OPT_COARSE_RBM, 0)
    xt, itt, _ = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12)
    assert cases.rel_err(xt, xd) < 1e-8 and itt < 0.8 * itj, (itt, itj)
    ctx.set_option(ob._lib.OPT_COARSE_RBM, 1)
    x2, it2, res2 = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12)
    assert cases.rel_err(x2, xd) < 1e-8 and cases.rel_err(xj, xd) < 1e-8
    assert np.all(x2[m.free_mask() == 0] == 0)
    assert it2 < itt < 0.8 * itj, (it2, itt, itj)
    x3, it3, _ = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12)
    assert it3 == it2
    np.testing.assert_array_equal(x3, x2)
    # a new K (other state): the coarse operator follows it
    U2 = cases.random_U(m, 0.004, seed=9)
    ref2 = oracle.Assembly(m).assemble(U2)
    ctx.set_U(U2)
    ctx.assemble()
    A2 = ref2.csr()
    xd2 = np.zeros(m.n_dofs)
    xd2[free] = spla.spsolve(A2[free][:, free].tocsc(), b[free])
    x4, it4, _ = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12)
    x5, it5, _ = ctx.pcg(b, ob.PRECOND_JACOBI, 1e-12)
    assert cases.rel_err(x5, xd2) < 1e-8 and cases.rel_err(x4, xd2) < 1e-8 and it4 < 0.8 * it5
    # one aggregate only (tiny mesh): still a valid preconditioner
    ms, _ = cases.box_model(3, 2, 2, mat="svk")
    cs = _ctx(ob, ms)
    cs.set_U(cases.random_U(ms, 0.01))
    cs.assemble()
    bs = rng.standard_normal(ms.n_dofs)
    refs = oracle.Assembly(ms).assemble(cases.random_U(ms, 0.01))
    xs, its, _ = cs.pcg(bs, ob.PRECOND_TWO_LEVEL, 1e-12)
    xds = np.zeros(ms.n_dofs)
    xds[ms.free_dofs] = spla.spsolve(refs.csr()[ms.free_dofs][:, ms.free_dofs].tocsc(), bs[ms.free_dofs])
    assert cases.rel_err(xs, xds) < 1e-8
    # only the streamed solver implements it
    ctx.set_option(ob._lib.OPT_CG_MODE, 1)
    with pytest.raises(ob.OnsasError):
        ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12)


def test_two_level_blocked_and_unblocked_coarse_inverse_agree(ob, oracle):
    """The coarse inverse by 12-row panels (default) and by single pivot rows are the same matrix up to rounding: same
    iteration count within one, same solution to 1e-10; sizes with full and partial last panels."""
    for grid in ((24, 12, 12), (30, 15, 15), (3, 2, 2)):
        m, _ = cases.box_model(*grid, mat="svk", jitter=0.1)
        U = cases.random_U(m, 0.002)
        b = np.random.default_rng(5).standard_normal(m.n_dofs)
        out = []
        for blocked in (1, 0):
            ctx = _ctx(ob, m)
            ctx.set_option(ob._lib.OPT_GJ_BLOCKED, blocked)
            ctx.set_U(U)
            ctx.assemble()
            out.append(ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12))
            ctx.close()
        (x1, it1, _), (x0, it0, _) = out
        assert abs(it1 - it0) <= 1, (grid, it1, it0)
        assert cases.rel_err(x1, x0) < 1e-9, grid


def test_newton_step_from_a_plain_c_program():
    """tests/abi_c/abi_smoke.c (strict C99, gcc, linked against the library): the reference's 6-tet uniaxial-extension cell through
    onsas_create ... onsas_newton_step ... onsas_destroy without Python or ctypes in between."""
    from tests.abi_c import run
    out = run.build_and_run()
    assert out.returncode == 0, out.stdout + out.stderr
    assert "one Newton step through the C ABI" in out.stdout, out.stdout


def test_coarse_operator_both_kernel_forms_agree():
    """E = Z^T (M K M) Z by k_coarse_assemble (own-aggregate blocks summed by all warps, the other targets dealt over thread
    groups) against the first, serial form of the kernel on the same K (ONSAS_COARSE_CHECK runs both and reports the largest
    difference): equal to rounding, E symmetric; with and without the rotations, jittered mesh, several aggregates."""
    import os
    import re
    import subprocess
    import sys
    code = (
        "import numpy as np, onsas_jl_b200 as ob\n"
        "from tests import cases\n"
        "from tests.test_gpu_parity import _ctx\n"
        "m, _ = cases.box_model(24, 12, 12, mat='neo', jitter=0.1)\n"
        "b = np.random.default_rng(8).standard_normal(m.n_dofs)\n"
        "for rbm in (1, 0):\n"
        "    ctx = _ctx(ob, m)\n"
        "    ctx.set_option(ob._lib.OPT_COARSE_RBM, rbm)\n"
        "    ctx.set_U(cases.random_U(m, 0.002))\n"
        "    ctx.assemble()\n"
        "    x, its, _ = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-10)\n"
        "    print('its', its)\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root,
                         env=dict(os.environ, ONSAS_COARSE_CHECK="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = re.findall(r"max \|E - E_other_form\| = (\S+), max \|E\| = (\S+), max \|E - E\^T\| = (\S+)", out.stderr)
    assert len(lines) == 2, out.stderr[-2000:]
    for d, e, a in lines:
        assert float(d) < 1e-12 * float(e) and float(a) < 1e-12 * float(e), lines


@pytest.mark.parametrize("dim", [1, 2])
def test_two_level_preconditioner_in_one_and_two_dimensions(ob, oracle, dim):
    """The coarse space in 1-D and 2-D (translations only: one or two coarse dofs per aggregate, 256 or 64 thread groups in
    k_coarse_assemble): a clamped chain and a braced, jittered truss grid pinned along one edge, several aggregates each -- the
    two-level solve equals the direct solve, needs no more iterations than Jacobi and is bitwise reproducible."""
    rng = np.random.default_rng(4)
    if dim == 1:
        m, _, _ = cases.clamped_truss(1500)   # 1501 nodes -> 4 aggregates
    else:
        nx, ny = 60, 40
        gx, gy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
        xyz = np.stack([gx.ravel(), gy.ravel()], axis=1).astype(float) + rng.uniform(-0.1, 0.1, ((nx + 1) * (ny + 1), 2))
        nid = lambda i, j: i * (ny + 1) + j
        bars = [(nid(i, j), nid(i + 1, j)) for i in range(nx) for j in range(ny + 1)]
        bars += [(nid(i, j), nid(i, j + 1)) for i in range(nx + 1) for j in range(ny)]
        bars += [(nid(i, j), nid(i + 1, j + 1)) for i in range(nx) for j in range(ny)]
        bars += [(nid(i + 1, j), nid(i, j + 1)) for i in range(nx) for j in range(ny)]
        bars = np.array(bars, np.int32)
        free = np.array([2 * n + c for n in range(len(xyz)) for c in range(2) if n >= ny + 1], np.int64)  # edge i = 0 pinned
        m = oracle.FlatModel(xyz=xyz, dim=2, trusses=bars, truss_area=rng.uniform(0.5, 1.5, len(bars)), truss_strain=1,
                             mat_kind=[0], mat_params=[[0.0, 50.0]], free_dofs=free)
    U = rng.uniform(-1e-3, 1e-3, m.n_dofs) * m.free_mask()
    ref = oracle.Assembly(m).assemble(U)
    import scipy.sparse.linalg as spla
    free = np.asarray(m.free_dofs)
    b = rng.standard_normal(m.n_dofs)
    xd = np.zeros(m.n_dofs)
    xd[free] = spla.spsolve(ref.csr()[free][:, free].tocsc(), b[free])
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    ctx.assemble()
    xj, itj, _ = ctx.pcg(b, ob.PRECOND_JACOBI, 1e-13)
    x2, it2, _ = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-13)
    assert cases.rel_err(xj, xd) < 1e-6 and cases.rel_err(x2, xd) < 1e-6, (cases.rel_err(xj, xd), cases.rel_err(x2, xd))
    assert it2 <= itj, (it2, itj)
    x3, it3, _ = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-13)
    assert it3 == it2
    np.testing.assert_array_equal(x3, x2)


def test_l2_prefetch_option_does_not_change_a_bit(ob):
    """ONSAS_OPT_CG_L2_PREFETCH (the producer warp asks for the slices behind its ring with cp.async.bulk.prefetch.L2): a pure
    cache hint -- same iterations, same bits, for the Jacobi and the two-level solver."""
    m, _ = cases.box_model(24, 12, 12, mat="neo", jitter=0.1)
    ctx = _ctx(ob, m)
    ctx.set_U(cases.random_U(m, 0.002))
    ctx.assemble()
    b = np.random.default_rng(3).standard_normal(m.n_dofs)
    for pre in (ob.PRECOND_JACOBI, ob.PRECOND_TWO_LEVEL):
        ctx.set_option(ob._lib.OPT_CG_L2_PREFETCH, 0)
        x0, it0, _ = ctx.pcg(b, pre, 1e-10)
        for pf in (1, 3, 64):
            ctx.set_option(ob._lib.OPT_CG_L2_PREFETCH, pf)
            x1, it1, _ = ctx.pcg(b, pre, 1e-10)
            assert it1 == it0
            np.testing.assert_array_equal(x1, x0)
    ctx.set_option(ob._lib.OPT_CG_L2_PREFETCH, 0)


@pytest.mark.parametrize("case", ["box", "random_numbering", "mixed"])
def test_assemble_host_is_bitwise_the_three_call_path(ob, oracle, case):
    """onsas_assemble_host (U in, F_int out, copies pipelined over slice ranges) against onsas_set_U + onsas_assemble +
    onsas_get_Fint: F_int, K, stress and strain bitwise equal for every chunk count, on a banded numbering, a random
    numbering (every chunk needs all of U) and a structure with two element families (accumulating second pass)."""
    if case == "box":
        m, _ = cases.box_model(20, 10, 10, mat="neo", jitter=0.1)
    elif case == "random_numbering":
        m0, _ = cases.box_model(11, 6, 5, mat="svk", jitter=0.15)
        rng = np.random.default_rng(11)
        n = m0.xyz.shape[0]
        perm = rng.permutation(n)
        inv = np.empty(n, np.int64)
        inv[perm] = np.arange(n)
        tets = perm[m0.tets][rng.permutation(len(m0.tets))].astype(np.int32)
        free = np.sort(perm[m0.free_dofs // 3] * 3 + m0.free_dofs % 3)
        m = oracle.FlatModel(xyz=m0.xyz[inv], tets=tets, mat_kind=m0.mat_kind, mat_params=m0.mat_params, free_dofs=free)
    else:
        m0, mesh = cases.box_model(7, 4, 4, jitter=0.1)
        rng = np.random.default_rng(2)
        bars = np.stack([np.arange(0, mesh.n_nodes - 1), np.arange(1, mesh.n_nodes)], axis=1).astype(np.int32)
        m = oracle.FlatModel(xyz=m0.xyz, tets=m0.tets, tet_mat=rng.integers(0, 3, len(m0.tets)), trusses=bars,
                             truss_mat=rng.integers(0, 2, len(bars)), truss_area=rng.uniform(0.1, 0.3, len(bars)),
                             truss_strain=1, mat_kind=[0, 1, 2], mat_params=[[0.58, 0.38], [0.83, 0.38], [1.0, 0.3]],
                             free_dofs=m0.free_dofs)
    U = cases.random_U(m, 0.03)
    ctx = _ctx(ob, m)
    ctx.set_U(U)
    ctx.assemble()
    F0, K0 = ctx.get_Fint(), ctx.get_csr()[2]
    S0 = ctx.get_stress_strain(ob.FAMILY_TET)
    ref = oracle.Assembly(m).assemble(U)
    assert cases.rel_err(F0, ref.F_int) < 1e-9
    for chunks, streams in ((4, 2), (4, 1), (1, 2), (3, 2), (64, 2), (7, 1)):
        ctx.set_option(ob._lib.OPT_HOST_CHUNKS, chunks)
        ctx.set_option(ob._lib.OPT_HOST_MID_WEIGHT, 1 if chunks == 3 else 3)
        ctx.set_option(ob._lib.OPT_HOST_STREAMS, streams)
        ctx.set_U(np.zeros_like(U))           # stale state that the call must replace
        ctx.assemble()
        F1 = ctx.assemble_host(U)
        np.testing.assert_array_equal(F1, F0)
        np.testing.assert_array_equal(ctx.get_csr()[2], K0)
        np.testing.assert_array_equal(ctx.get_U(), U)
        S1 = ctx.get_stress_strain(ob.FAMILY_TET)
        np.testing.assert_array_equal(S1[0], S0[0])
        np.testing.assert_array_equal(S1[1], S0[1])
    # the state it leaves behind feeds the solver like onsas_assemble's does
    b = np.random.default_rng(4).standard_normal(m.n_dofs)
    x1 = ctx.pcg(b, ob.PRECOND_JACOBI, 1e-12)[0]
    ctx.set_U(U)
    ctx.assemble()
    np.testing.assert_array_equal(ctx.pcg(b, ob.PRECOND_JACOBI, 1e-12)[0], x1)
    ctx.close()


def test_library_side_reordering_is_invisible_to_the_caller(ob, oracle):
    """ONSAS_OPT_REORDER = 1 on a randomly numbered mesh (nodes AND elements permuted, as a Gmsh mesh arrives,
    Interfaces/Gmsh.jl:25-83): the library renumbers along a Z-curve internally; every result comes back in the caller's
    numbering -- K, F_int, stress / strain bitwise equal to the un-reordered context (entries are summed in element order
    either way), solves equal to the solver tolerance, device-side loads equal."""
    m, mesh = cases.box_model(9, 6, 5, mat="neo", jitter=0.1)
    rng = np.random.default_rng(4)
    perm, eperm = rng.permutation(mesh.n_nodes), rng.permutation(mesh.n_tets)
    xyz = np.empty_like(m.xyz)
    xyz[perm] = m.xyz
    tets = perm[m.tets[eperm]].astype(np.int32)
    free = np.sort(perm[m.free_dofs // 3] * 3 + m.free_dofs % 3)
    kw = dict(tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    plain = ob.context_from_flat(xyz, **kw)
    reord = ob.context_from_flat(xyz, reorder=1, **kw)
    gm = oracle.FlatModel(xyz=xyz, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    U = cases.random_U(gm, 0.02)
    for c in (plain, reord):
        c.set_U(U)
        c.assemble()
    np.testing.assert_array_equal(reord.get_Fint(), plain.get_Fint())
    for a, b in zip(plain.get_csr(), reord.get_csr()):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(plain.get_stress_strain(), reord.get_stress_strain()):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(reord.assemble_host(U), plain.get_Fint())
    ref = oracle.Assembly(gm).assemble(U)
    assert cases.rel_err(reord.get_Fint(), ref.F_int) < 1e-12
    b = rng.standard_normal(gm.n_dofs)
    x0, it0, _ = plain.pcg(b, ob.PRECOND_JACOBI, 1e-12)
    for pre in (ob.PRECOND_JACOBI, ob.PRECOND_TWO_LEVEL):
        x1, it1, _ = reord.pcg(b, pre, 1e-12)
        assert np.abs(x1 - x0).max() < 1e-9 * np.abs(x0).max()
    faces = perm[mesh.faces["x1"]].astype(np.int32)
    for c in (plain, reord):
        c.add_face_load(faces, 0, [-1.0, 0.0, 0.0])
        c.apply_loads([0.05])
        c.set_U(np.zeros(gm.n_dofs))
    np.testing.assert_array_equal(reord.get_Fext(), plain.get_Fext())
    for _ in range(8):
        i0, i1 = plain.newton_step(ob.PRECOND_JACOBI, 1e-13), reord.newton_step(ob.PRECOND_JACOBI, 1e-13)
    assert cases.rel_err(reord.get_U(), plain.get_U()) < 1e-8 and i1.norm_r < 1e-8 * i1.norm_Fext
    st0, st1 = plain.table_stats(), reord.table_stats()
    assert st1["nnz_blocks"] == st0["nnz_blocks"] and st1["padded_block_slots"] <= st0["padded_block_slots"]
    for c in (plain, reord):
        c.close()


def test_non_finite_residual_stops_the_solve_at_once(ob):
    """A NaN in F_ext (or K, U) makes the CG residual non-finite: `res <= tol` is never true, so without a guard the solve would
    run to maxiter = n_free iterations.  The solvers stop at once and report ONSAS_ERR_BREAKDOWN; the context stays usable."""
    import time
    m, mesh = cases.box_model(8, 4, 4, mat="svk")
    for mode in (0, 1, 2):
        ctx = ob.context_from_flat(m.xyz, tets=m.tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=m.free_dofs)
        ctx.set_option(ob._lib.OPT_CG_MODE, mode)
        F = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (0.1, 0.0, 0.0))
        bad = F.copy()
        bad[m.free_dofs[5]] = np.nan
        ctx.set_Fext(bad)
        t0 = time.perf_counter()
        with pytest.raises(ob.OnsasError) as ei:
            ctx.newton_step(ob.PRECOND_JACOBI)
        assert ei.value.status == ob._lib.ERR_BREAKDOWN and time.perf_counter() - t0 < 5.0
        ctx.set_U(np.zeros(mesh.n_nodes * 3))
        ctx.set_Fext(F)
        info = ctx.newton_step(ob.PRECOND_JACOBI)                 # the same context solves again
        assert info.cg_residual <= info.cg_tol and np.isfinite(ctx.get_U()).all()
        ctx.close()


def test_two_contexts_with_different_meshes_interleave(ob):
    """The dynamic shared-memory limit of a kernel is a per-device property shared by all contexts: a second context with a
    narrower mesh (a 1-D truss chain plans a few KB for the streamed CG, a tet mesh ~200 KB) must not lower it under the
    first one (round-1 advisor finding)."""
    m, mesh = cases.box_model(8, 4, 4, mat="svk")
    big = ob.context_from_flat(m.xyz, tets=m.tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=m.free_dofs)
    F = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (0.1, 0.0, 0.0))
    big.set_Fext(F)
    i0 = big.newton_step(ob.PRECOND_JACOBI)
    mt, fext, p = cases.clamped_truss(50)
    small = ob.context_from_flat(mt.xyz, trusses=mt.trusses, truss_area=mt.truss_area, truss_strain=ob.STRAIN_GREEN,
                                 mat_kind=mt.mat_kind, mat_params=mt.mat_params, free_dofs=mt.free_dofs)
    small.set_Fext(fext(0.1))
    small.newton_step(ob.PRECOND_JACOBI, 1e-12)
    big.set_U(np.zeros(mesh.n_nodes * 3))
    i1 = big.newton_step(ob.PRECOND_JACOBI)                       # would fail with "invalid argument" if the limit had been lowered
    assert i1.cg_iters == i0.cg_iters and i1.norm_r == i0.norm_r
    for pre in (ob.PRECOND_TWO_LEVEL,):
        big.set_U(np.zeros(mesh.n_nodes * 3))
        small.set_U(np.zeros(len(mt.xyz)))
        big.newton_step(pre)
        small.newton_step(pre, 1e-12)
        big.newton_step(pre)
    big.close()
    small.close()


def test_assemble_host_graph_replay_is_bitwise_the_eager_pipeline(ob):
    """onsas_assemble_host with the same PINNED buffers call after call: the second call captures the pipeline (copies, slice-range
    kernels, events on four streams) into a CUDA graph, later calls replay it.  Every call -- eager, capturing, replayed, with new
    contents in the same buffers, after an option change (the graph is dropped and re-captured) -- gives bitwise the three-call
    result; pageable buffers and ONSAS_OPT_HOST_GRAPH = 0 stay on the eager path."""
    import torch
    m, mesh = cases.box_model(24, 12, 12, mat="neo")
    ctx = ob.context_from_flat(m.xyz, tets=m.tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=m.free_dofs)
    n = mesh.n_nodes * 3
    hU = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    hF = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    rng = np.random.default_rng(8)
    ctx.set_option(ob._lib.OPT_HOST_GRAPH, 1)

    def three_calls(U):
        ctx.set_U(U)
        ctx.assemble()
        return ctx.get_Fint()

    for trial in range(6):
        if trial == 4:
            ctx.set_option(ob._lib.OPT_HOST_CHUNKS, 5)            # drops the captured graph: a new one is captured for 5 ranges
        U = 0.02 * rng.standard_normal(n)
        hU[:] = U
        hF[:] = np.nan
        ctx.assemble_host(hU, hF)
        ref = three_calls(U)
        np.testing.assert_array_equal(hF, ref)
        s1, e1 = ctx.get_stress_strain()
    ctx.set_option(ob._lib.OPT_HOST_GRAPH, 0)
    hF[:] = np.nan
    ctx.assemble_host(hU, hF)
    np.testing.assert_array_equal(hF, ref)
    np.testing.assert_array_equal(ctx.assemble_host(np.array(hU)), ref)   # pageable buffers: eager
    ctx.close()


def test_material_swap_on_a_finalized_mesh_keeps_the_state(ob):
    """onsas_set_materials after onsas_finalize_mesh (replace!(s, material), Structures.jl) re-derives the element kinds in
    place: U, F_ext and the load patterns survive, and the next assembly equals a fresh context of the new material bit for bit
    (round-1 advisor finding: the swap used to invalidate the mesh and silently drop the solver state)."""
    m, mesh = cases.box_model(8, 4, 4, mat="svk")
    K, mu = 1.0 / (3 * 0.4), 1.0 / 2.6
    ctx = ob.context_from_flat(m.xyz, tets=m.tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=m.free_dofs)
    ctx.add_face_load(mesh.faces["x1"], 0, [1.0, 0.0, 0.0])
    ctx.apply_loads([0.3])
    U = cases.random_U(m, 0.02)
    ctx.set_U(U)
    F0 = ctx.get_Fext()
    ctx.set_materials([ob.MAT_NEOHOOKEAN], [[K, mu]])
    np.testing.assert_array_equal(ctx.get_U(), U)
    np.testing.assert_array_equal(ctx.get_Fext(), F0)
    ctx.apply_loads([0.6])                                         # the pattern is still registered
    np.testing.assert_allclose(ctx.get_Fext(), 2 * F0, rtol=1e-15)
    ctx.assemble()
    fresh = ob.context_from_flat(m.xyz, tets=m.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[K, mu]], free_dofs=m.free_dofs)
    fresh.set_U(U)
    fresh.assemble()
    np.testing.assert_array_equal(ctx.get_Fint(), fresh.get_Fint())
    for a, b in zip(ctx.get_csr(), fresh.get_csr()):
        np.testing.assert_array_equal(a, b)
    info = ctx.newton_step(ob.PRECOND_TWO_LEVEL)                   # and the coarse operator follows the new K
    assert info.cg_residual <= info.cg_tol
    with pytest.raises(ob.OnsasError):
        ctx.set_materials([7], [[1.0, 1.0]])                       # unknown kind
    ctx.close()
    fresh.close()
