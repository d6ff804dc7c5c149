"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise).
  * one process per GPU under torchrun (tests/multi_gpu_check.py): assembly, SpMV, PCG in all solver modes and a 9-step
    Newton solve against the oracle, numpy and native partitioner;
  * ONE process driving both GPUs behind the C ABI (onsas_create_multi): the whole interface with global vectors in the
    caller's numbering, against a single-device context of the same structure and bitwise against the torchrun run."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

from onsas_jl_b200 import meshgen as mg
from tests import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")


def test_two_gpu_distributed_parity(tmp_path, ob, oracle):
    _need_two_gpus()
    dump = str(tmp_path / "U_torchrun.npy")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, ONSAS_MULTI_DUMP=dump))
    print(out.stdout[-2000:])
    assert "MULTI_GPU_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    # the same 9-step Newton solve from ONE process driving both devices: bitwise the torchrun result
    m2, mesh2 = cases.box_model(12, 6, 6, mat="neo")
    ctx = ob.context_from_flat(m2.xyz, tets=m2.tets, mat_kind=m2.mat_kind, mat_params=m2.mat_params, free_dofs=m2.free_dofs, device=[0, 1])
    Fo = mg.global_face_load(mesh2.n_nodes, mesh2.xyz, mesh2.faces["x1"], (-1.0, 0.0, 0.0))
    tols = oracle.ConvergenceSettings(1e-10, 1e-10, 20)
    ctx.set_U(np.zeros(mesh2.n_nodes * 3))
    for t in np.linspace(1 / 9, 1.0, 9):
        ctx.set_Fext(Fo * t)
        dU_rel = dr_rel = 1e12
        it = 0
        while oracle.criterion(dU_rel, dr_rel, it, tols) == "NotConvergedYet":
            info = ctx.newton_step(ob.PRECOND_JACOBI, 1e-13)
            dU_rel = info.norm_dU / info.norm_U if info.norm_U > 0 else math.inf
            dr_rel = info.norm_r / info.norm_Fext
            it += 1
    np.testing.assert_array_equal(ctx.get_U().reshape(-1, 3), np.load(dump))
    ctx.close()


def test_multi_device_context_one_process(ob, oracle):
    """onsas_create_multi: every entry point of the C ABI with global arrays, one host thread, two devices -- against the
    single-device context of the same structure (K, F_int, records bitwise; solves to the north-star tolerances) and the oracle."""
    _need_two_gpus()
    m, mesh = cases.box_model(10, 5, 4, mat="svk", jitter=0.1)
    rng = np.random.default_rng(11)
    perm = rng.permutation(mesh.n_nodes)                          # arbitrary caller numbering
    xyz = np.empty_like(m.xyz)
    xyz[perm] = m.xyz
    tets = perm[m.tets].astype(np.int32)
    free = np.sort(perm[m.free_dofs // 3] * 3 + m.free_dofs % 3)
    mk = dict(tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    one = ob.context_from_flat(xyz, device=0, **mk)
    two = ob.context_from_flat(xyz, device=[0, 1], **mk)
    assert two._lib.onsas_device_count(two._h) == 2 and one._lib.onsas_device_count(one._h) == 1
    gm = oracle.FlatModel(xyz=xyz, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    U = cases.random_U(gm, 0.02)
    for c in (one, two):
        c.set_U(U)
        c.assemble()
    np.testing.assert_array_equal(two.get_U(), U)
    np.testing.assert_array_equal(two.get_Fint(), one.get_Fint())             # rows are summed in the same order on any partition
    r1, c1, v1 = one.get_csr()
    r2, c2, v2 = two.get_csr()
    np.testing.assert_array_equal(r1, r2)
    np.testing.assert_array_equal(c1, c2)
    np.testing.assert_array_equal(v1, v2)
    for a, b in zip(one.get_stress_strain(), two.get_stress_strain()):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(one.eval_elements(first=7, count=300), two.eval_elements(first=7, count=300)):
        np.testing.assert_array_equal(a, b)
    F1 = np.array(one.assemble_host(U))
    np.testing.assert_array_equal(two.assemble_host(U), F1)
    ref = oracle.Assembly(gm).assemble(U)
    assert cases.rel_err(two.get_Fint(), ref.F_int) < 1e-12
    # SpMV / PCG with global vectors
    mask = gm.free_mask()
    x = rng.standard_normal(gm.n_dofs) * mask
    assert cases.rel_err(two.spmv(x), one.spmv(x)) < 1e-14
    b = rng.standard_normal(gm.n_dofs)
    x1, it1, _ = one.pcg(b, ob.PRECOND_JACOBI, 1e-12)
    for mode in (0, 2):
        two.set_option(ob._lib.OPT_CG_MODE, mode)
        x2, it2, _ = two.pcg(b, ob.PRECOND_JACOBI, 1e-12)
        assert np.abs(x2 - x1).max() < 1e-9 * np.abs(x1).max() and abs(it2 - it1) <= 2
    two.set_option(ob._lib.OPT_CG_MODE, 0)
    with pytest.raises(ob.OnsasError):
        two.set_option(ob._lib.OPT_CG_MODE, 1)                    # the per-phase NCCL solver needs one process per GPU
    x3, it3, _ = two.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-12)
    assert np.abs(x3 - x1).max() < 1e-8 * np.abs(x1).max()
    # device-side loads: faces given with global node ids
    faces = perm[mesh.faces["x1"]].astype(np.int32)
    for c in (one, two):
        c.add_face_load(faces, 0, [1.0, 0.0, 0.0])
        c.add_face_load(faces, 1, [1.0])
        c.add_nodal_load(perm[mesh.node_sets["x0"][:5]], [0.0, 2.0, 0.0])
        c.apply_loads([0.3, -0.2, 0.1])
    np.testing.assert_array_equal(two.get_Fext(), one.get_Fext())
    # Newton iterations: same counts, same state
    for c in (one, two):
        c.set_U(np.zeros(gm.n_dofs))
    for _ in range(7):
        i1 = one.newton_step(ob.PRECOND_JACOBI, 1e-13)
        i2 = two.newton_step(ob.PRECOND_JACOBI, 1e-13)
        assert i2.norm_r == pytest.approx(i1.norm_r, rel=1e-6, abs=1e-11) and i2.norm_Fext == pytest.approx(i1.norm_Fext, rel=1e-13)
    assert cases.rel_err(two.get_U(), one.get_U()) < 1e-9 and i2.norm_r < 1e-8 * i2.norm_Fext
    assert cases.rel_err(two.get_dU(), one.get_dU()) < 1e-6 or np.abs(two.get_dU()).max() < 1e-12
    st = two.table_stats()
    assert st["nnz_blocks"] == one.table_stats()["nnz_blocks"]
    for c in (one, two):
        c.close()


def test_reference_api_on_two_devices(ob):
    """The mirrored reference API with `NewtonRaphson(device=[0, 1])`: examples/uniaxial_extension at 20 x 10 x 10 cells,
    iteration counts of the shipped example and the analytic end state."""
    _need_two_gpus()
    mesh = mg.box_tet_mesh(20, 10, 10, 2.0, 1.0, 1.0)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (3.0, 0.0, 0.0))
    s = ob.Structure.from_arrays(mesh.xyz, tets=mesh.tets, materials=[ob.SVK(E=1.0, nu=0.3)], free_dofs=free, fext=lambda t: unit * t)
    sol = ob.solve(ob.NonLinearStaticAnalysis(s, NSTEPS=8),
                   ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 30), preconditioner="two_level", cg_reltol=1e-11, device=[0, 1]))
    assert sol.iterations() == [6, 5, 5, 4, 4, 4, 5, 5]
    Ua = mg.homogeneous_field(mesh.xyz, 2.0, math.sqrt(0.1))
    assert np.abs(sol.U[-1] - Ua).max() < 1e-8 * np.abs(Ua).max()
    assert np.abs(sol.tet_stress[-1][:, 0] - 3.0).max() < 1e-7


def test_truss_lattice_on_two_devices_and_reordered(ob):
    """Trusses through the multi-device context (and through the library-side renumbering on one device): a clamped, pulled and
    sheared braced lattice, large-displacement Newton -- same iterates as the plain single-device context."""
    _need_two_gpus()
    lat = mg.truss_lattice(10, 4, 4, 2.0)
    E, A = 210e9, 2.5e-3
    fixed = {c: lat.node_sets["x0"] for c in range(3)}
    free = mg.free_dofs_from_fixed(lat.n_nodes, 3, fixed)
    F = np.zeros((lat.n_nodes, 3))
    F[lat.node_sets["x1"], 0] = 0.02 * E * A
    F[lat.node_sets["x1"], 2] = 0.002 * E * A
    kw = dict(trusses=lat.bars, truss_area=np.full(lat.n_bars, A), truss_strain=ob.STRAIN_GREEN, mat_kind=[ob.MAT_SVK],
              mat_params=[[0.0, E / 2]], free_dofs=free)
    ctxs = [ob.context_from_flat(lat.xyz, device=0, **kw), ob.context_from_flat(lat.xyz, device=[0, 1], **kw),
            ob.context_from_flat(lat.xyz, device=0, reorder=1, **kw), ob.context_from_flat(lat.xyz, device=[1, 0], reorder=2, **kw)]
    for c in ctxs:
        c.set_Fext(F.ravel())
    for it in range(7):
        infos = [c.newton_step(ob.PRECOND_JACOBI, 1e-13) for c in ctxs]
        for i in infos[1:]:
            assert i.norm_r == pytest.approx(infos[0].norm_r, rel=1e-6, abs=1e-9 * infos[0].norm_Fext)
    U0 = ctxs[0].get_U()
    assert infos[0].norm_r < 1e-8 * infos[0].norm_Fext and np.abs(U0).max() > 0.01
    for c in ctxs[1:]:
        assert cases.rel_err(c.get_U(), U0) < 1e-9
        for a, b in zip(c.get_stress_strain(ob.FAMILY_TRUSS), ctxs[0].get_stress_strain(ob.FAMILY_TRUSS)):
            np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-6 * E * 1e-3)
    for c in ctxs:
        c.close()
    # a 1-D chain (block size 1) on two devices: examples/clamped_truss, Green strain
    mt, fext, p = cases.clamped_truss(100)
    kw1 = dict(trusses=mt.trusses, truss_area=mt.truss_area, truss_strain=ob.STRAIN_GREEN, mat_kind=mt.mat_kind, mat_params=mt.mat_params,
               free_dofs=mt.free_dofs)
    a, b = ob.context_from_flat(mt.xyz, device=0, **kw1), ob.context_from_flat(mt.xyz, device=[0, 1], **kw1)
    for c in (a, b):
        c.set_Fext(fext(0.3))
        for _ in range(6):
            info = c.newton_step(ob.PRECOND_JACOBI, 1e-13, cg_maxiter=5000)
    assert cases.rel_err(b.get_U(), a.get_U()) < 1e-9 and info.norm_r < 1e-8 * info.norm_Fext
    a.close()
    b.close()
