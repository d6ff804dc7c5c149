"""The BASELINE.json configs at full size on one B200: size-independent parity properties + timings.
usage: python scripts/config_sweep.py [c2 c3 c4 c5]   (default: all)
Prints one JSON line per config."""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402
from onsas_jl_b200 import meshgen as mg  # noqa: E402

L = ob._lib
stream = torch.cuda.Stream()


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def tet_config(name, mesh, kind, params, free, U, Fext, analytic_U=None, newton=True, cg_reltol=None):
    t0 = time.time()
    ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[kind], mat_params=[params], free_dofs=free)
    t_fin = time.time() - t0
    ctx.set_stream(stream.cuda_stream)
    ctx.set_U(U)
    ctx.set_Fext(Fext)
    ms_asm = timed(ctx.assemble, 10)
    ctx.synchronize()
    Fint = ctx.get_Fint()
    mask = np.zeros(mesh.n_nodes * 3, bool)
    mask[free] = True
    res = float(np.abs((Fext - Fint)[mask]).max() / max(np.abs(Fext).max(), 1e-300))
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(U.size) * mask, rng.standard_normal(U.size) * mask
    Kx, Ky = ctx.spmv(x), ctx.spmv(y)
    sym = float(abs(y @ Kx - x @ Ky) / abs(y @ Kx))
    ms_spmv = timed(ctx.spmv_resident, 20)
    st = ctx.table_stats()
    out = {"config": name, "n_tets": int(mesh.n_tets), "n_dofs": int(mesh.n_nodes * 3), "finalize_s": round(t_fin, 2),
           "ms_assemble": ms_asm, "tets_per_s": mesh.n_tets / ms_asm * 1e3, "algo_GBps": 1600 * mesh.n_tets / ms_asm / 1e6,
           "residual_at_state": res, "K_symmetry": sym, "ms_spmv": ms_spmv,
           "spmv_GBps": (76 * st["nnz_blocks"] + 17 * mesh.n_nodes * 3) / ms_spmv / 1e6, "nnz_blocks": st["nnz_blocks"]}
    if newton:
        info = ctx.newton_step(ob.PRECOND_JACOBI, cg_reltol)
        out.update({"newton_ms_assemble": info.ms_assemble, "newton_ms_solve": info.ms_solve, "cg_iters": int(info.cg_iters),
                    "us_per_cg_iter": 1e3 * info.ms_solve / max(int(info.cg_iters), 1),
                    "rel_residual_in": info.norm_r / max(info.norm_Fext, 1e-300)})
        if analytic_U is not None:
            Un = ctx.get_U()
            out["U_err_vs_analytic"] = float(np.abs(Un - analytic_U).max() / np.abs(analytic_U).max())
    ctx.close()
    print(json.dumps(out), flush=True)


def c2():
    mesh, free, U_half, U_prev, Fext = bench.build_problem(55, 1)
    a1, b1 = bench._neo_state(0.5)
    ctx_U = mg.homogeneous_field(mesh.xyz, a1, b1)
    # start the Newton step from a slightly perturbed analytic state: one step must land back on it
    rng = np.random.default_rng(1)
    mask = np.zeros(mesh.n_nodes * 3)
    mask[free] = 1
    tet_config("c2 uniaxial_compression NeoHookean cube 55^3", mesh, ob.MAT_NEOHOOKEAN, (bench.KBULK, bench.MU), free, ctx_U, Fext,
               analytic_U=ctx_U)


def c3():
    Ri, Re, Lz, E, nu, p = 100.0, 200.0, 30.0, 210.0, 0.3, 10.0
    mesh = mg.cylinder_tet_mesh(48, 576, 30, Ri, Re, Lz)
    fixed = {2: mesh.node_sets["z_caps"], 0: mesh.node_sets["outer_on_y_axis"], 1: mesh.node_sets["outer_on_x_axis"]}
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, fixed)
    Fp = mg.pressure_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["inner"], p)
    r = np.linalg.norm(mesh.xyz[:, :2], axis=1)
    A = (1 + nu) * (1 - 2 * nu) * Ri ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    B = (1 + nu) * Ri ** 2 * Re ** 2 * p / (E * (Re ** 2 - Ri ** 2))
    Ulame = np.zeros((mesh.n_nodes, 3))
    Ulame[:, :2] = mesh.xyz[:, :2] / r[:, None] * (A * r + B / r)[:, None]
    # linear material: one Newton step from U = 0 is the linear solve; compare with the Lame field (mesh-dependent error)
    tet_config("c3 cylinder_internal_pressure IsotropicLinearElastic (48,576,30)", mesh, ob.MAT_ISOLINEAR, (E, nu), free,
               np.zeros(mesh.n_nodes * 3), Fp, analytic_U=Ulame.ravel(), cg_reltol=1e-8)


def c4(n=188):
    mesh = mg.box_tet_mesh(n, n, n, 1.0, 1.0, 1.0)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    E, nu, p = 1.0, 0.3, 3.0
    lam, G = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    U = mg.homogeneous_field(mesh.xyz, 2.0, math.sqrt(0.1))       # analytic end state of uniaxial_extension (alpha = 2)
    Fext = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (p, 0.0, 0.0))
    tet_config(f"c4 SVK cube {n}^3", mesh, ob.MAT_SVK, (lam, G), free, U, Fext, analytic_U=U, newton=(n <= 188))


def c5(n=112):
    mesh = mg.truss_lattice(n, n, n, 2.0)
    E, A = 210e9, 2.5e-3
    nn = mesh.n_nodes
    free = np.arange(nn * 3, dtype=np.int64)
    t0 = time.time()
    ctx = ob.context_from_flat(mesh.xyz, trusses=mesh.bars, truss_area=np.full(mesh.n_bars, A), truss_strain=ob.STRAIN_GREEN,
                               mat_kind=[ob.MAT_SVK], mat_params=[[0.0, E / 2]], free_dofs=free)
    t_fin = time.time() - t0
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option(L.OPT_TRUSS_MINBLOCKS, int(os.environ.get("ONSAS_TRUSS_MINB", "3")))
    eps = 1e-3
    U = np.zeros((nn, 3))
    U[:, 0] = eps * mesh.xyz[:, 0]                                   # homogeneous stretch: interior nodes stay in equilibrium
    ctx.set_U(U.ravel())
    ms_asm = timed(ctx.assemble, 10)
    ctx.synchronize()
    Fint = ctx.get_Fint().reshape(-1, 3)
    g = np.rint(mesh.xyz / 2.0).astype(int)
    interior = np.all((g > 0) & (g < n), axis=1)
    f_scale = E * A * eps
    s, e = ctx.get_stress_strain(ob.FAMILY_TRUSS)
    d = mesh.xyz[mesh.bars[:, 1]] - mesh.xyz[mesh.bars[:, 0]]
    l0 = np.linalg.norm(d, axis=1)
    l1 = np.linalg.norm(d * np.array([1 + eps, 1, 1]), axis=1)
    eg = (l1 ** 2 - l0 ** 2) / (2 * l0 ** 2)
    ctx.set_U(np.zeros(nn * 3))
    ctx.assemble()
    t = np.zeros((nn, 3))
    t[:, 1] = 1.0
    Kt = ctx.spmv(t.ravel())                                        # rigid translation is in the null space at U = 0
    ms_spmv = timed(ctx.spmv_resident, 20)
    print(json.dumps({"config": f"c5 truss lattice {n}^3 Green strain", "n_bars": int(mesh.n_bars), "n_dofs": int(nn * 3),
                      "finalize_s": round(t_fin, 2), "ms_assemble": ms_asm, "bars_per_s": mesh.n_bars / ms_asm * 1e3,
                      "algo_GBps": 456 * mesh.n_bars / ms_asm / 1e6,
                      "interior_equilibrium": float(np.abs(Fint[interior]).max() / f_scale),
                      "strain_err": float(np.abs(e[:, 0] - eg).max() / eps), "stress_err": float(np.abs(s[:, 0] - E * eg * l1 / l0).max() / (E * eps)),
                      "rigid_translation_K_t": float(np.abs(Kt).max() / (E * A / 2.0)), "ms_spmv": ms_spmv}), flush=True)
    ctx.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3", "c4", "c5"]
    for w in which:
        try:
            {"c2": c2, "c3": c3, "c4": c4, "c5": c5}[w]()
        except Exception as ex:  # keep going: one config must not hide the others
            print(json.dumps({"config": w, "error": repr(ex)[:500]}), flush=True)
