"""CPU checks of the DEVICE-side arithmetic (onsas.jl_b200/csrc/element_math.cuh compiled with g++) and of the
host-built tables, by walking the kernels' control flow serially (tests/hostsim).  The real kernels are checked
on the GPU in test_gpu_*.py; this catches arithmetic / indexing mistakes before a GPU box is spent on them."""
import numpy as np
import pytest

from tests import cases


@pytest.mark.parametrize("mat", ["svk", "neo", "iso"])
def test_element_rows_match_oracle(oracle, hostsim, mat):
    m, U = cases.random_tet_model(2000, mat)
    f, K, s, e = oracle.eval_tets(m, U)
    f2, K2, s2, e2 = hostsim.eval_tets(m, U)
    # north star: per-element f_int and K_t within 1e-10 relative
    for a, b in ((f2, f), (K2, K), (s2, s), (e2, e)):
        scale = np.abs(b).max(axis=1, keepdims=True)
        assert (np.abs(a - b) / scale).max() < 1e-10


@pytest.mark.parametrize("strain", [0, 1])
@pytest.mark.parametrize("dim", [1, 2, 3])
def test_truss_rows_match_oracle(oracle, hostsim, strain, dim):
    rng = np.random.default_rng(11)
    n = 300
    xyz = rng.uniform(0, 1, (2 * n, dim))
    xyz[1::2] += 0.5
    bars = np.arange(2 * n, dtype=np.int32).reshape(n, 2)
    m = oracle.FlatModel(xyz=xyz, dim=dim, trusses=bars, truss_area=rng.uniform(0.5, 2, n), truss_strain=strain,
                         mat_kind=[0, 1], mat_params=[[0.4, 1.1], [2.0, 0.7]], truss_mat=rng.integers(0, 2, n),
                         free_dofs=np.arange(2 * n * dim))
    U = rng.uniform(-0.2, 0.2, 2 * n * dim)
    f, K, s, e = oracle.eval_trusses(m, U)
    f2, K2, s2, e2 = hostsim.eval_trusses(m, U)
    for a, b in ((f2, f), (K2, K), (s2, s), (e2, e)):
        assert cases.rel_err(a, b) < 1e-12


@pytest.mark.parametrize("mat,grid", [("svk", (5, 3, 4)), ("neo", (3, 3, 3)), ("iso", (9, 2, 1))])
def test_row_owner_assembly_matches_reference_order(oracle, hostsim, mat, grid):
    m, _ = cases.box_model(*grid, mat=mat, jitter=0.15)
    U = cases.random_U(m)
    ref = oracle.Assembly(m).assemble(U)
    sim = hostsim.HostSim(m)
    out = [sim.assemble(U, threads=t) for t in (192, 64, 256)]
    rp, ci, v, Fi, to, _ = out[0]
    np.testing.assert_array_equal(rp, ref.rowptr)
    np.testing.assert_array_equal(ci, ref.col)
    row_scale = np.repeat(np.maximum.reduceat(np.abs(ref.val), ref.rowptr[:-1]), np.diff(ref.rowptr))
    assert (np.abs(v - ref.val) / row_scale).max() < 1e-13
    assert cases.rel_err(Fi, ref.F_int) < 1e-13
    assert cases.rel_err(to[:, :9], ref.tet_sig) < 1e-13
    for o in out[1:]:  # the result does not depend on the launch configuration
        np.testing.assert_array_equal(o[2], v)
        np.testing.assert_array_equal(o[3], Fi)


def test_mixed_materials_and_families(oracle, hostsim):
    """tets of three materials + trusses sharing nodes in one structure (ACCUM path of the second family)."""
    m, mesh = cases.box_model(3, 2, 2, jitter=0.1)
    rng = np.random.default_rng(2)
    bars = np.stack([np.arange(0, mesh.n_nodes - 1), np.arange(1, mesh.n_nodes)], axis=1).astype(np.int32)
    mm = oracle.FlatModel(xyz=m.xyz, tets=m.tets, tet_mat=rng.integers(0, 3, len(m.tets)), trusses=bars,
                          truss_mat=rng.integers(0, 2, len(bars)), truss_area=rng.uniform(0.1, 0.3, len(bars)),
                          truss_strain=1, mat_kind=[0, 1, 2], mat_params=[[0.58, 0.38], [0.83, 0.38], [1.0, 0.3]],
                          free_dofs=m.free_dofs)
    U = cases.random_U(mm, 0.03)
    ref = oracle.Assembly(mm).assemble(U)
    rp, ci, v, Fi, to, tr = hostsim.HostSim(mm).assemble(U)
    np.testing.assert_array_equal(ci, ref.col)
    assert cases.rel_err(v, ref.val) < 1e-13 and cases.rel_err(Fi, ref.F_int) < 1e-13
    assert cases.rel_err(tr[:, 0], ref.truss_sig[:, 0]) < 1e-13


def test_partial_ownership_rows(oracle, hostsim):
    """Multi-GPU layout: only the first n_owned nodes are rows; halo nodes appear as columns only."""
    m, _ = cases.box_model(4, 2, 2, jitter=0.1)
    U = cases.random_U(m)
    ref = oracle.Assembly(m).assemble(U).csr()
    n_own = 20
    sim = hostsim.HostSim(m, n_rows=n_own)
    rp, ci, v, Fi, _, _ = sim.assemble(U)
    import scipy.sparse as sp
    got = sp.csr_matrix((v, ci, rp), shape=(3 * n_own, m.n_dofs)).toarray()
    np.testing.assert_allclose(got, ref[:3 * n_own].toarray(), rtol=1e-13, atol=1e-15)


def test_bsell_spmv_and_pcg_phase_order(oracle, hostsim):
    m, _ = cases.box_model(5, 3, 3, jitter=0.1)
    ref = oracle.Assembly(m).assemble(cases.random_U(m, 0.02))
    sim = hostsim.HostSim(m)
    sim.assemble(cases.random_U(m, 0.02))
    mask = m.free_mask()
    rng = np.random.default_rng(4)
    x = rng.standard_normal(m.n_dofs) * mask
    assert cases.rel_err(sim.spmv(mask, x), (ref.csr() @ x) * mask) < 1e-13
    b = rng.standard_normal(m.n_dofs)
    d = np.where(mask, ref.csr().diagonal(), 1.0)
    for pre, diag in ((0, None), (1, d)):
        xs, its, _ = sim.pcg(mask, b, pre, 1e-10)
        xo, ito, _ = oracle.cg(ref.rowptr, ref.col, ref.val, mask, b, diag=diag, reltol=1e-10)
        assert abs(its - ito) <= 2 and cases.rel_err(xs, xo) < 1e-8


def test_table_limits(hostsim, oracle):
    m, _ = cases.box_model(2, 2, 2)
    st = hostsim.HostSim(m).stats()
    assert st[0] == -(-m.n_nodes // 8) and st[2] <= st[1] and st[3] <= 8 * 24
    bad = oracle.FlatModel(xyz=m.xyz, tets=np.array([[0, 1, 2, 999]], np.int32), free_dofs=[0])
    with pytest.raises(ValueError):
        hostsim.HostSim(bad)


def test_host_range_plan_of_the_pipelined_assembly(hostsim, oracle):
    """onsas_assemble_host sends U as growing prefixes: range k may start once nodes [0, node_hi[k]) have arrived.  The plan
    must cover every slice once, be monotone, and node_hi[k] must bound every node touched by the elements that the
    slices of range k evaluate (all elements incident to their rows); a banded numbering pipelines (node_hi grows with
    k), a random numbering needs (almost) all of U at once."""
    m, mesh = cases.box_model(12, 6, 6, mat="svk")
    rng = np.random.default_rng(5)
    n = m.xyz.shape[0]
    perm = rng.permutation(n)
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    shuffled = oracle.FlatModel(xyz=m.xyz[inv], tets=perm[m.tets].astype(np.int32), mat_kind=m.mat_kind, mat_params=m.mat_params,
                                free_dofs=np.sort(perm[m.free_dofs // 3] * 3 + m.free_dofs % 3))
    for model, banded in ((m, True), (shuffled, False)):
        sim = hostsim.HostSim(model)
        n_slices = sim.stats()[0]
        tets = np.asarray(model.tets)
        for chunks, weight in ((1, 1), (2, 3), (4, 4), (12, 3), (1000, 2)):
            s0, hi = sim.host_plan(chunks, weight)
            assert s0[0] == 0 and s0[-1] == n_slices and np.all(np.diff(s0) >= 0) and len(hi) == len(s0) - 1
            assert len(hi) == min(chunks, n_slices) and np.all(np.diff(hi) >= 0) and hi[-1] <= n
            for k in range(len(hi)):
                rows = np.arange(8 * s0[k], min(8 * s0[k + 1], n))
                if len(rows) == 0:
                    continue
                touched = tets[np.isin(tets, rows).any(axis=1)]            # elements incident to the range's rows
                need = max(int(touched.max()) + 1 if len(touched) else 0, int(rows.max()) + 1)
                assert hi[k] >= need and (k > 0 or hi[k] == need)
            if chunks == 12:
                inner = np.diff(s0)[1:-1]
                assert np.diff(s0)[0] <= inner.min() and np.diff(s0)[-1] <= inner.min() + 1   # short first / last range
                if banded:
                    assert hi[0] < 0.35 * n          # the first kernel starts after a small prefix of U
                else:
                    assert hi[0] > 0.9 * n           # random numbering: (almost) everything first


def test_slice_node_lists_the_assembly_kernel_stages(hostsim, oracle):
    """The per-slice node lists (tables.cpp, round 2): every list is strictly ascending, every pair's 16-bit indices map back to
    its element's nodes, exactly one pair of every element carries the "writes the element record" bit -- also when only part of
    the nodes are owned (multi-GPU: the pair of the element's first OWNED node), and on a structured mesh a slice lists a small
    neighbourhood (the point of staging it in shared memory) while a random numbering lists almost every node a pair touches."""
    m, mesh = cases.box_model(10, 6, 5, mat="svk")
    sim = hostsim.HostSim(m)
    chk = sim.check_slice_nodes(0, m.tets)
    assert chk["violations"] == 0 and chk["max_list"] <= 100          # 8 consecutive x-nodes +- 1 in every direction: <= 10 x 3 x 3
    rng = np.random.default_rng(2)
    perm = rng.permutation(m.xyz.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    shuffled = oracle.FlatModel(xyz=m.xyz[inv], tets=perm[m.tets].astype(np.int32), mat_kind=m.mat_kind, mat_params=m.mat_params,
                                free_dofs=np.arange(len(perm) * 3))
    chk2 = hostsim.HostSim(shuffled).check_slice_nodes(0, shuffled.tets)
    assert chk2["violations"] == 0 and chk2["entries"] > 1.5 * chk["entries"]
    part = hostsim.HostSim(m, n_rows=200).check_slice_nodes(0, m.tets)       # only the first 200 nodes are rows: elements whose
    assert part["violations"] == 0 and 0 < part["entries"] < chk["entries"]   # first node is a halo node still have one writer
    mt, _, _ = cases.von_mises_truss(0)
    assert hostsim.HostSim(mt).check_slice_nodes(1, mt.trusses)["violations"] == 0


def test_host_built_tables_match_their_recorded_hashes(hostsim, oracle):
    """The table layout IS the contract between tables.cpp and the kernels (pair lists, slice node lists, contribution codes,
    slice headers, the BSELL pattern): an FNV-1a hash over every array, for structured and random numbering, trusses in 1-3 D,
    both families together, partial ownership and a 400-bar hub, must equal the recorded one -- whatever the thread count.
    A deliberate layout change regenerates tests/golden/table_hashes.json with scripts/table_hashes.py."""
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("table_hashes", os.path.join(root, "scripts", "table_hashes.py"))
    th = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(th)
    golden = json.load(open(os.path.join(root, "tests", "golden", "table_hashes.json")))
    seen = {}
    for name, (m, n_rows) in th.meshes().items():
        sim = hostsim.HostSim(m, n_rows=n_rows)
        seen[name] = {"hash": f"{sim.tables_hash():016x}", "stats": sim.stats()}
    assert seen == golden
