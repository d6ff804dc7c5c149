"""CPU tests of the host-side logic: the mirror of the reference API (flattening, dof numbering, free dofs, load
lumping, convergence bookkeeping), the C-ABI library (loads, exports every declared symbol, refuses to run
without a GPU) and the multi-GPU partition / halo plan."""
import copy
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- C ABI

def test_library_exports_every_declared_symbol(ob):
    ob.build()
    hdr = open(os.path.join(ROOT, "include", "onsas_cuda.h")).read()
    declared = set(re.findall(r"\b(onsas_[a-z_A-Z0-9]+)\s*\(", hdr))
    declared -= {"onsas_ctx"}
    lib = C.CDLL(ob._lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/onsas_cuda.h but not exported"
    assert declared == set(ob._lib.SIGNATURES), "python binding and header disagree"
    assert lib.onsas_version() >= 100


def test_option_keys_status_codes_and_enums_match_the_header(ob):
    """Every ONSAS_OPT_* / ONSAS_ERR_* / ONSAS_PRECOND_* / ONSAS_MAT_* constant of include/onsas_cuda.h has the same value in the
    Python binding (a caller that passes the binding's constant reaches the option the header documents), option keys are unique,
    and setting an unknown key is an error rather than a silent no-op (checked without a device: the context cannot be created
    here, so the key table is compared, not exercised)."""
    hdr = open(os.path.join(ROOT, "include", "onsas_cuda.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(ONSAS_[A-Z0-9_]+)\s+(-?\d+)\b", hdr)}
    opts = {k: v for k, v in defs.items() if k.startswith("ONSAS_OPT_")}
    assert len(opts) >= 18 and len(set(opts.values())) == len(opts), "duplicate option key"
    L = ob._lib
    checked = 0
    for name, value in defs.items():
        short = name[len("ONSAS_"):]
        if hasattr(L, short):
            assert getattr(L, short) == value, (name, value, getattr(L, short))
            checked += 1
    for name in opts:
        assert hasattr(L, name[len("ONSAS_"):]), f"{name} has no constant in the Python binding"
    assert checked >= len(opts) + 3


def test_header_is_plain_c_and_the_library_links_from_c(ob):
    """include/onsas_cuda.h compiled as strict C99 (-pedantic -Werror) by gcc, linked against libonsas_cuda.so the way a ccall /
    cgo / JNI binding would: no C++ type, no torch symbol in the interface.  Without a device the C program sees the loud failure
    the header promises (ONSAS_ERR_CUDA, no context, a message); with one it runs a Newton step (see the -m gpu twin)."""
    from tests.abi_c import run
    out = run.build_and_run()
    assert out.returncode == 0, out.stdout + out.stderr
    assert "C linkage ok" in out.stdout or "one Newton step through the C ABI" in out.stdout


def test_no_cpu_fallback(ob):
    """Without a CUDA device the product path must fail loudly (this test is skipped on a GPU box)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ob.OnsasError) as ei:
        ob.DeviceContext(0)
    assert ei.value.status == ob._lib.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "onsas.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "element_math.cuh", f
    src = open(os.path.join(pkg, "csrc", "element_math.cuh")).read()
    assert "#include \"../../oracle" not in src and "liboracle" not in src


# ----------------------------------------------------------------------------- convergence bookkeeping

def test_residuals_iteration_step_semantics(ob):
    """test/structural_solvers/structural_solvers.jl:26-70."""
    tols = ob.ConvergenceSettings(1e-5, 1e-3, 100)  # positional order U, force, iter
    assert tols.rel_res_force_tol == 1e-3 and tols.rel_U_tol == 1e-5 and tols.max_iter == 100
    r = ob.ResidualsIterationStep()
    assert r.iter == 0 and isinstance(r.criterion, ob.NotConvergedYet)
    r.reset()
    assert r.ΔU_rel >= 1e3 and r.Δr_rel >= 1e3
    assert isinstance(ob.isconverged(r, tols), ob.NotConvergedYet)
    dU = np.linalg.norm([1e-10, 1e-10])
    r.update(dU, dU / np.linalg.norm([1e-1, 1e-4]), dU, dU / np.linalg.norm([1e3, 1e3]))
    assert r.iter == 1
    assert not isinstance(ob.isconverged(r, tols), ob.NotConvergedYet)
    r.reset()
    assert r.iter == 0 and isinstance(ob.isconverged(r, tols), ob.NotConvergedYet)


def test_isconverged_order_and_assertions(ob):
    """StructuralSolvers.jl:148-171: residual first, then dU, then iter > max_iter; zero norms are errors."""
    tols = ob.ConvergenceSettings(1e-6, 1e-6, 3)
    r = ob.ResidualsIterationStep()
    r.update(1.0, 1e-9, 1.0, 1e-9)
    assert isinstance(ob.isconverged(r, tols), ob.ResidualForceCriterion)
    r.update(1.0, 1e-9, 1.0, 1.0)
    assert isinstance(ob.isconverged(r, tols), ob.DeltaUCriterion)
    r.update(1.0, math.inf, 1.0, 1.0)   # first Newton iteration: ||U|| = 0 -> Inf
    assert isinstance(ob.isconverged(r, tols), ob.NotConvergedYet)
    r.update(1.0, 1.0, 1.0, 1.0)
    with pytest.warns(UserWarning):
        assert isinstance(ob.isconverged(r, tols), ob.MaxIterCriterion)  # iter = 4 > 3
    r.update(0.0, 0.0, 1.0, 1.0)
    with pytest.raises(AssertionError):
        ob.isconverged(r, tols)


# ----------------------------------------------------------------------------- model mirror

def _uniaxial_structure(ob, p=3.0):
    """examples/uniaxial_extension/uniaxial_extension.jl:36-110 built through the mirrored API."""
    Lx, Ly, Lz = 2.0, 1.0, 1.0
    n = [ob.Node(0.0, 0.0, 0.0), ob.Node(0.0, 0.0, Lz), ob.Node(0.0, Ly, Lz), ob.Node(0.0, Ly, 0.0),
         ob.Node(Lx, 0.0, 0.0), ob.Node(Lx, 0.0, Lz), ob.Node(Lx, Ly, Lz), ob.Node(Lx, Ly, 0.0)]
    n1, n2, n3, n4, n5, n6, n7, n8 = n
    m = ob.Mesh(nodes=n)
    f = [ob.TriangularFace(n5, n8, n6), ob.TriangularFace(n6, n8, n7), ob.TriangularFace(n4, n1, n2),
         ob.TriangularFace(n4, n2, n3), ob.TriangularFace(n6, n2, n1), ob.TriangularFace(n6, n1, n5),
         ob.TriangularFace(n1, n4, n5), ob.TriangularFace(n4, n8, n5)]
    m.faces += f
    t = [ob.Tetrahedron(n1, n4, n2, n6), ob.Tetrahedron(n6, n2, n3, n4), ob.Tetrahedron(n4, n3, n6, n7),
         ob.Tetrahedron(n4, n1, n5, n6), ob.Tetrahedron(n4, n6, n5, n8), ob.Tetrahedron(n4, n7, n6, n8)]
    m.elements += t
    ob.set_dofs(m, "u", 3)
    svk = ob.SVK(E=1.0, nu=0.3, label="svk")
    mat = ob.StructuralMaterial((svk, t))
    bcs = ob.StructuralBoundaryCondition((ob.FixedField("u", [1]), f[2:4]), (ob.FixedField("u", [2]), f[4:6]),
                                         (ob.FixedField("u", [3]), f[6:8]),
                                         (ob.GlobalLoad("u", lambda tt: [p * tt, 0, 0]), f[0:2]))
    return ob.Structure(m, mat, bcs), n, t


def test_structure_flattening(ob, oracle):
    s, n, t = _uniaxial_structure(ob)
    fl = s.flat
    assert [nd.dofs["u"] for nd in n[:2]] == [[1, 2, 3], [4, 5, 6]]           # Meshes.jl:85-98
    assert fl.tets.tolist()[0] == [0, 3, 1, 5] and fl.n_dofs == 24
    assert len(fl.free_dofs) == 12 and s.num_free_dofs == 12
    # free dofs keep node order with the fixed ones removed (Structures.jl:129-142)
    assert np.all(np.diff(fl.free_dofs) > 0)
    # face load: p*A/3 per face node, duplicates summed (GlobalLoadBoundaryConditions.jl:50-68)
    F = fl.fext(0.5).reshape(-1, 3)
    assert F[:, 0].sum() == pytest.approx(3.0 * 0.5 * 1.0) and np.all(F[:, 1:] == 0)
    np.testing.assert_allclose(F[[4, 5, 6, 7], 0], 1.5 * np.array([1, 2, 1, 2]) / 6)
    # same numbers as the array-based generator
    from onsas_jl_b200 import meshgen as mg
    mesh = mg.box_tet_mesh(1, 1, 1)
    Fa = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (1.5, 0, 0))
    assert Fa.sum() == pytest.approx(F.sum())
    # the flat model drives the oracle to the analytic end state (alpha = 2)
    fm = oracle.FlatModel(xyz=fl.xyz, tets=fl.tets, mat_kind=fl.mat_kind, mat_params=fl.mat_params, free_dofs=fl.free_dofs)
    res = oracle.newton_solve(fm, np.linspace(1 / 8, 1, 8), fl.fext, oracle.ConvergenceSettings(1e-8, 1e-8, 30))
    assert 1 + res.U[-1][3 * 6] / 2.0 == pytest.approx(2.0, rel=1e-4)


def test_boundary_condition_lumping(ob):
    """test/boundary_conditions/boundary_conditions.jl:62-105."""
    n1, n2, n3, n4 = ob.Node(0, 0, 0), ob.Node(0, 1, 0), ob.Node(0, 0, 1), ob.Node(2, 0, 1)
    face = ob.TriangularFace(n1, n2, n3)
    tet = ob.Tetrahedron(n1, n2, n3, n4)
    mesh = ob.Mesh(nodes=[n1, n2, n3, n4], elements=[tet], faces=[face])
    ob.set_dofs(mesh, "u", 3)
    t = 2.0
    vals = lambda tt: [math.sin(tt), tt, tt ** 2]  # noqa: E731
    mat = ob.StructuralMaterial((ob.SVK(1.0, 1.0), [tet]))

    def fext(bc, ent):
        s = ob.Structure(mesh, mat, ob.StructuralBoundaryCondition((bc, [ent])))
        return s.flat.fext(t)

    np.testing.assert_allclose(fext(ob.GlobalLoad("u", vals), n1)[:3], vals(t))
    A = face.area()
    assert A == 0.5
    np.testing.assert_allclose(fext(ob.GlobalLoad("u", vals), face)[:9], np.tile(np.array(vals(t)) * A / 3, 3))
    np.testing.assert_allclose(fext(ob.GlobalLoad("u", vals), tet), np.tile(np.array(vals(t)) * (2 / 6) / 4, 4))
    # pressure acts along -n: the face lies in x = 0 with normal +x
    np.testing.assert_allclose(fext(ob.Pressure("u", lambda tt: tt ** 2), face)[:9], -np.tile([t ** 2 * A / 3, 0, 0], 3))


def test_materials_and_cross_sections(ob):
    """test/materials/materials.jl:86-107, test/cross_sections."""
    svk = ob.SVK(E=210e9, nu=0.3)
    lam, G = svk.lame_parameters()
    assert svk.elasticity_modulus() == pytest.approx(210e9) and svk.poisson_ratio() == pytest.approx(0.3)
    neo = ob.NeoHookean(E=210e9, nu=0.3)
    assert neo.bulk_modulus() == pytest.approx(lam + 2 * G / 3) and neo.elasticity_modulus() == pytest.approx(210e9)
    iso = ob.IsotropicLinearElastic(lam=lam, G=G)
    assert iso.E == pytest.approx(210e9) and iso.nu == pytest.approx(0.3)
    assert ob.Circle(2.0).area() == pytest.approx(math.pi) and ob.Square(3.0).area() == 9.0
    assert ob.Rectangle(2.0, 3.0).area() == 6.0


def test_analysis_load_factors_and_deepcopy(ob):
    s, _, _ = _uniaxial_structure(ob)
    sa = ob.NonLinearStaticAnalysis(s, NSTEPS=8)
    np.testing.assert_allclose(sa.load_factors(), np.linspace(1 / 8, 1.0, 8))  # LinRange(t1/N, t1, N)
    with pytest.raises(ValueError):
        ob.NonLinearStaticAnalysis(s, NSTEPS=8, initial_step=9)
    sa._ctx = object()
    sb = copy.deepcopy(sa)   # solve() deep-copies the analysis; the device handle must not travel
    assert sb._ctx is None and sb.s is sa.s
    assert ob.NewtonRaphson().cg_reltol == pytest.approx(math.sqrt(np.finfo(float).eps))
    assert ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 30)).tol.max_iter == 30


def test_unsupported_models_are_rejected(ob):
    n1, n2 = ob.Node(0.0), ob.Node(1.0)
    tr = ob.Truss(n1, n2, ob.Square(1.0), ob.GreenStrain)
    mesh = ob.Mesh(nodes=[n1, n2], elements=[tr])
    ob.set_dofs(mesh, "u", 1)
    with pytest.raises(TypeError):  # Trusses.jl dispatches on AbstractHyperElasticMaterial only
        ob.Structure(mesh, ob.StructuralMaterial((ob.IsotropicLinearElastic(1.0, 0.3), [tr])),
                     ob.StructuralBoundaryCondition())
    with pytest.raises(ValueError):
        ob.set_dofs(mesh, "u", 1)  # "Dof symbol u already exists." (Meshes.jl:88-90)


# ----------------------------------------------------------------------------- synthetic meshes

def test_box_mesh_matches_reference_split():
    from onsas_jl_b200 import meshgen as mg
    m = mg.box_tet_mesh(3, 2, 2)
    X = m.xyz[m.tets]
    J = np.stack([X[:, 0] - X[:, 1], X[:, 3] - X[:, 1], X[:, 2] - X[:, 1]], axis=2)
    vol = np.linalg.det(J) / 6
    assert np.all(vol > 0) and vol.sum() == pytest.approx(2.0)
    # conforming: every interior face is shared by exactly two tets
    faces = np.sort(np.concatenate([m.tets[:, [0, 1, 2]], m.tets[:, [0, 1, 3]], m.tets[:, [0, 2, 3]], m.tets[:, [1, 2, 3]]]), axis=1)
    _, cnt = np.unique(faces, axis=0, return_counts=True)
    assert set(cnt) == {1, 2} and (cnt == 1).sum() == 2 * 2 * (3 * 2 + 3 * 2 + 2 * 2)
    assert mg.face_areas(m.xyz, m.faces["x1"]).sum() == pytest.approx(1.0)
    # the single-hex mesh is the reference's own 6 tets (uniaxial_extension.jl:45-72) up to node numbering
    m1 = mg.box_tet_mesh(1, 1, 1)
    ref_nodes = np.array([(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0), (2, 0, 0), (2, 0, 1), (2, 1, 1), (2, 1, 0)], float)
    ref_tets = np.array([(1, 4, 2, 6), (6, 2, 3, 4), (4, 3, 6, 7), (4, 1, 5, 6), (4, 6, 5, 8), (4, 7, 6, 8)]) - 1
    np.testing.assert_array_equal(m1.xyz[m1.tets], ref_nodes[ref_tets])


def test_cylinder_and_lattice_meshes():
    from onsas_jl_b200 import meshgen as mg
    c = mg.cylinder_tet_mesh(3, 16, 2)
    X = c.xyz[c.tets]
    J = np.stack([X[:, 0] - X[:, 1], X[:, 3] - X[:, 1], X[:, 2] - X[:, 1]], axis=2)
    vol = np.linalg.det(J) / 6
    assert np.all(vol > 0)
    assert vol.sum() == pytest.approx(np.pi * (200 ** 2 - 100 ** 2) * 30, rel=0.03)
    tri = c.faces["inner"]
    a, b, cc = c.xyz[tri[:, 0]], c.xyz[tri[:, 1]], c.xyz[tri[:, 2]]
    nrm = np.cross(b - a, cc - a)
    cen = (a + b + cc) / 3
    assert np.all((nrm[:, :2] * cen[:, :2]).sum(axis=1) < 0)   # normals point to the axis
    assert np.all(c.xyz[c.node_sets["outer_on_y_axis"], 0] == 0) and np.all(c.xyz[c.node_sets["outer_on_x_axis"], 1] == 0)
    t = mg.truss_lattice(3, 2, 2)
    assert t.n_bars == len(np.unique(np.sort(t.bars, axis=1), axis=0)) and np.all(t.bars[:, 0] != t.bars[:, 1])


# ----------------------------------------------------------------------------- partition / halo plan

def test_rcb_partition_and_halo_plan(oracle, hostsim):
    from onsas_jl_b200 import partition as pt
    m, mesh = cases.box_model(6, 3, 3, jitter=0.1)
    order, ranges = pt.rcb_order(m.xyz, 4)
    assert sorted(order.tolist()) == list(range(m.n_nodes)) and ranges[-1] == m.n_nodes
    assert max(np.diff(ranges)) - min(np.diff(ranges)) <= 1
    xyz, tets, inv = pt.renumber(order, m.xyz, m.tets)
    free = np.sort((inv[m.free_dofs // 3] * 3 + m.free_dofs % 3))
    parts = [pt.build_local_part(r, ranges, xyz, tets=tets, free_dofs=free) for r in range(4)]
    assert sum(p.n_owned for p in parts) == m.n_nodes and sum(int((p.free_dofs < 3 * p.n_owned).sum()) for p in parts) == len(free)
    for p in parts:
        # what I send to k is exactly what k expects from me, in the same (global id) order
        for k, r in enumerate(p.nbr_rank):
            q = parts[r]
            j = list(q.nbr_rank).index(p.rank)
            mine = p.local_to_global[p.send_nodes[p.send_ptr[k]:p.send_ptr[k + 1]]]
            theirs = q.local_to_global[q.n_owned + q.recv_ptr[j]: q.n_owned + q.recv_ptr[j + 1]]
            np.testing.assert_array_equal(mine, theirs)
    # owned rows assembled locally == the same rows of the global matrix (bitwise: same summation order)
    gm = oracle.FlatModel(xyz=xyz, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    U = cases.random_U(gm)
    gsim = hostsim.HostSim(gm)
    grp, gci, gv, gF, _, _ = gsim.assemble(U)
    for p in parts:
        lm = oracle.FlatModel(xyz=p.xyz, tets=p.tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=p.free_dofs)
        rp, ci, v, Fi, _, _ = hostsim.HostSim(lm, n_rows=p.n_owned).assemble(p.scatter_global(U, 3))
        dofs = p.owned_global_dofs(3)
        np.testing.assert_array_equal(Fi, gF[dofs])
        for lr, gr in ((0, dofs[0]), (len(dofs) - 1, dofs[-1])):
            lcols = (p.local_to_global[ci[rp[lr]:rp[lr + 1]] // 3] * 3 + ci[rp[lr]:rp[lr + 1]] % 3)
            o = np.argsort(lcols)
            np.testing.assert_array_equal(lcols[o], gci[grp[gr]:grp[gr + 1]])
            np.testing.assert_array_equal(v[rp[lr]:rp[lr + 1]][o], gv[grp[gr]:grp[gr + 1]])


def _gloo_worker(rank, world, port, tmp):
    import torch.distributed as dist
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from onsas_jl_b200 import partition as pt
    from oracle import oracle as O
    from tests.hostsim import hostsim_py as H
    m, _ = cases.box_model(6, 3, 3, jitter=0.1)
    order, ranges = pt.rcb_order(m.xyz, world)
    xyz, tets, inv = pt.renumber(order, m.xyz, m.tets)
    free = np.sort((inv[m.free_dofs // 3] * 3 + m.free_dofs % 3))
    p = pt.build_local_part(rank, ranges, xyz, tets=tets, free_dofs=free)
    lm = O.FlatModel(xyz=p.xyz, tets=p.tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=p.free_dofs)
    sim = H.HostSim(lm, n_rows=p.n_owned)
    gm = O.FlatModel(xyz=xyz, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    U = cases.random_U(gm)
    sim.assemble(p.scatter_global(U, 3))
    # distributed y = K x with the halo plan (the exchange libonsas_cuda does with NCCL send/recv)
    xg = np.random.default_rng(9).standard_normal(gm.n_dofs) * gm.free_mask()
    xl = np.zeros(p.n_local * 3)
    xl[:p.n_owned * 3] = xg[p.owned_global_dofs(3)]
    reqs, bufs = [], []
    for k, r in enumerate(p.nbr_rank):
        sn = p.send_nodes[p.send_ptr[k]:p.send_ptr[k + 1]]
        sb = torch.from_numpy(xl.reshape(-1, 3)[sn].copy())
        rb = torch.zeros((int(p.recv_ptr[k + 1] - p.recv_ptr[k]), 3), dtype=torch.float64)
        reqs += [dist.isend(sb, int(r)), dist.irecv(rb, int(r))]
        bufs.append((k, rb, sb))
    for q in reqs:
        q.wait()
    for k, rb, _ in bufs:
        xl.reshape(-1, 3)[p.n_owned + p.recv_ptr[k]: p.n_owned + p.recv_ptr[k + 1]] = rb.numpy()
    mask = np.zeros(p.n_local * 3, np.uint8)
    mask[p.free_dofs] = 1
    yl = sim.spmv(mask, xl)
    # global dot product through an all-reduce (what the PCG does with ncclAllReduce)
    d = torch.tensor([float(xl[:p.n_owned * 3] @ yl)], dtype=torch.float64)
    dist.all_reduce(d)
    np.save(os.path.join(tmp, f"y{rank}.npy"), yl)
    np.save(os.path.join(tmp, f"d{rank}.npy"), d.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_halo_exchange_gloo(oracle, hostsim, tmp_path):
    import torch.multiprocessing as mp
    from onsas_jl_b200 import partition as pt
    world, port = 2, 29541 + os.getpid() % 500
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    m, _ = cases.box_model(6, 3, 3, jitter=0.1)
    order, ranges = pt.rcb_order(m.xyz, world)
    xyz, tets, inv = pt.renumber(order, m.xyz, m.tets)
    free = np.sort((inv[m.free_dofs // 3] * 3 + m.free_dofs % 3))
    gm = oracle.FlatModel(xyz=xyz, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    A = oracle.Assembly(gm).assemble(cases.random_U(gm)).csr()
    xg = np.random.default_rng(9).standard_normal(gm.n_dofs) * gm.free_mask()
    yref = (A @ xg) * gm.free_mask()
    y = np.concatenate([np.load(tmp_path / f"y{r}.npy") for r in range(world)])
    assert cases.rel_err(y, yref) < 1e-13
    d0, d1 = np.load(tmp_path / "d0.npy"), np.load(tmp_path / "d1.npy")
    assert d0 == d1 and d0[0] == pytest.approx(xg @ yref, rel=1e-12)


# ----------------------------------------------------------------------------- result hand-back

def test_write_vtk_from_flat_solution(ob, tmp_path):
    """write_vtk (Interfaces/VTK.jl:209-262) from the flat per-step arrays: cell types, point vector field with the
    reference's component names, one scalar cell array per tensor label taken at the (i, j) the label names."""
    from onsas_jl_b200 import vtk
    s, n, t = _uniaxial_structure(ob)
    sa = ob.NonLinearStaticAnalysis(s, np.linspace(0.5, 1.0, 2))
    sol = ob.Solution(sa, ob.NewtonRaphson())
    rng = np.random.default_rng(3)
    fl = s.flat
    for _ in range(2):
        sol.U.append(rng.standard_normal(fl.n_dofs))
        sol.tet_stress.append(rng.standard_normal((len(fl.tets), 9)))
        sol.tet_strain.append(rng.standard_normal((len(fl.tets), 9)))
    path = ob.write_vtk(sol, str(tmp_path / "uniaxial"), 2)
    assert path.endswith("uniaxial.vtu") and os.path.exists(path)
    arrays, meta = vtk.read_vtu(path)
    assert meta["n_points"] == 8 and meta["n_cells"] == 6
    np.testing.assert_array_equal(arrays["Points"], fl.xyz)
    np.testing.assert_array_equal(arrays["connectivity"].reshape(-1, 4), fl.tets)
    np.testing.assert_array_equal(arrays["offsets"], 4 * np.arange(1, 7))
    assert set(arrays["types"].tolist()) == {10}                                   # VTK_TETRA
    np.testing.assert_array_equal(arrays["Displacement"], sol.U[1].reshape(-1, 3))
    assert meta["component_names"]["Displacement"] == ["ux", "uy", "uz"]
    sig = sol.tet_stress[1].reshape(-1, 3, 3).transpose(0, 2, 1)                   # column-major 3x3 -> [e, i, j]
    eps = sol.tet_strain[1].reshape(-1, 3, 3).transpose(0, 2, 1)
    np.testing.assert_array_equal(arrays["σxx"], sig[:, 0, 0])
    np.testing.assert_array_equal(arrays["τyz"], sig[:, 1, 2])
    np.testing.assert_array_equal(arrays["τzy"], sig[:, 2, 1])
    np.testing.assert_array_equal(arrays["τzx"], sig[:, 2, 0])
    np.testing.assert_array_equal(arrays["γxy"], eps[:, 0, 1])
    np.testing.assert_array_equal(arrays["γyx"], eps[:, 1, 0])
    assert len([k for k in arrays if k[0] in "στϵγ"]) == 18
    # only displacements
    arrays_u, _ = vtk.read_vtu(ob.write_vtk(sol, str(tmp_path / "u_only.vtu"), 1, fields=("u",)))
    assert "σxx" not in arrays_u and "Displacement" in arrays_u
    # the time series: one file per stored step + the ParaView collection with the load factors as times
    pvd = ob.write_vtk(sol, str(tmp_path / "series"))
    text = open(pvd, encoding="utf-8").read()
    assert pvd.endswith("series.pvd") and 'timestep="0.5"' in text and 'timestep="1.0"' in text
    assert os.path.exists(tmp_path / "series_timestep_1.vtu") and os.path.exists(tmp_path / "series_timestep_2.vtu")
    assert 'file="series_timestep_2.vtu"' in text
    # NaN in a cell field is refused like the reference's @assert (VTK.jl:139-140); a bad step index too
    sol.tet_stress[0][0, 0] = np.nan
    with pytest.raises(AssertionError):
        ob.write_vtk(sol, str(tmp_path / "bad"), 1)
    with pytest.raises(IndexError):
        ob.write_vtk(sol, str(tmp_path / "bad"), 3)
    # trusses become VTK_LINE cells after the tets
    p2 = vtk.write_vtu(str(tmp_path / "mixed"), np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), trusses=[[0, 1], [1, 2]],
                       cell_data={"σxx": np.array([1.0, 2.0])})
    a2, m2 = vtk.read_vtu(p2)
    assert m2["n_cells"] == 2 and a2["types"].tolist() == [3, 3] and a2["offsets"].tolist() == [2, 4]
    assert a2["Points"].shape == (3, 3) and np.all(a2["Points"][:, 2] == 0)


# ----------------------------------------------------------------------------- bench.py contract (reference arm runs on CPU)

def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the restated CPU algorithm on the host cores) prints exactly one JSON line with the
    contract's keys; the other ranks of a torchrun launch print nothing."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "6", "--steps", "2", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="0")).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["metric"] == "tet_fint_Kt_assembled_elements_per_s" and d["unit"] == "tets/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    silent = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1"))
    assert silent.returncode == 0 and silent.stdout.strip() == ""


# ----------------------------------------------------------------------------- documentation stays in step with the tree

def test_design_md_cites_tests_and_profiles_that_exist():
    """Every `test_…` name and every `profiles/…` file that DESIGN.md / README.md cite exists in the tree."""
    import glob
    text = open(os.path.join(ROOT, "DESIGN.md"), encoding="utf-8").read() + open(os.path.join(ROOT, "README.md"), encoding="utf-8").read()
    defined = set()
    for path in glob.glob(os.path.join(ROOT, "tests", "test_*.py")):
        defined |= set(re.findall(r"^def (test_\w+)", open(path, encoding="utf-8").read(), flags=re.M))
        defined.add(os.path.basename(path)[:-3])
    cited = set(re.findall(r"`(?:tests/)?(?:test_\w+\.py::)?(test_\w+?)(?:\[[^\]]*\])?(?:\.py)?`", text))
    missing = sorted(t for t in cited if t not in defined and not any(d.startswith(t.rstrip("_…")) for d in defined))
    assert not missing, missing
    for rel in set(re.findall(r"`(profiles/r\d+/[\w./-]+\.\w+)`", text)):
        assert os.path.exists(os.path.join(ROOT, rel)), rel
    for rel in set(re.findall(r"`(profiles/r\d+)`", text)):
        assert os.path.isdir(os.path.join(ROOT, rel)), rel
