"""BASELINE.json configs[2] (cylinder, ~5 M tets, 1-2 GPUs) and configs[4] (10 M-bar space-truss lattice, large-displacement
Newton on 8 GPUs) at full size on N GPUs, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/config_multi.py c3|c5
(or plain `python scripts/config_multi.py c3` on one GPU).  Prints one JSON line (rank 0) with timings and the
size-independent parity properties each config offers."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import onsas_jl_b200 as ob  # noqa: E402
from onsas_jl_b200 import meshgen as mg  # noqa: E402

world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


def allmax(*v):
    t = torch.tensor(v, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def allsum(*v):
    t = torch.tensor(v, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def make(xyz, kind, params, free, tets=None, trusses=None, area=None, strain=0):
    if world == 1:
        return ob.context_from_flat(xyz, tets=tets, trusses=trusses, truss_area=area, truss_strain=strain, mat_kind=kind, mat_params=params,
                                    free_dofs=free, device=local_rank), None, len(xyz)
    from onsas_jl_b200 import multigpu
    P = ob.NativePartition(xyz, world, tets=tets, trusses=trusses, truss_area=area, free_dofs=free)
    ctx = multigpu.make_distributed_context(P, kind, params, dist, local_rank, truss_strain=strain)
    l2g = P.local_to_global(rank).astype(np.int64)
    n_own = P.sizes(rank)["n_owned"]
    P.close()
    return ctx, l2g, n_own


def timed_assembly(ctx, stream, reps=10):
    for _ in range(3):
        ctx.assemble()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    e0.record(stream)
    for _ in range(reps):
        ctx.assemble()
    e1.record(stream)
    torch.cuda.synchronize()
    return allmax(e0.elapsed_time(e1) / reps)[0]


def c3():
    """examples/cylinder_internal_pressure scaled to (48, 576, 30) = 4 976 640 tets, IsotropicLinearElastic: the Newton step
    from U = 0 IS the linear solve; radial displacement vs the plane-strain Lame field, a second step finds no residual."""
    Ri, Re, Lz, E, nu, p = 100.0, 200.0, 30.0, 210.0, 0.3, 10.0
    mesh = mg.cylinder_tet_mesh(48, 576, 30, Ri, Re, Lz)
    fixed = {2: mesh.node_sets["z_caps"], 0: mesh.node_sets["outer_on_y_axis"], 1: mesh.node_sets["outer_on_x_axis"]}
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, fixed)
    Fp = mg.pressure_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["inner"], p)
    ctx, l2g, n_own = make(mesh.xyz, [ob.MAT_ISOLINEAR], [[E, nu]], free, tets=mesh.tets)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    loc = (lambda v: v) if l2g is None else (lambda v: v.reshape(-1, 3)[l2g].ravel())
    ctx.set_Fext(loc(Fp))
    ctx.set_U(loc(np.zeros(mesh.n_nodes * 3)))
    ms_asm = timed_assembly(ctx, stream)
    out = {"config": "configs[2] cylinder_internal_pressure IsotropicLinearElastic (48,576,30)", "n_gpus": world, "n_tets": int(mesh.n_tets),
           "n_dofs": int(mesh.n_nodes * 3), "assembly_ms": ms_asm, "tets_per_s": mesh.n_tets / ms_asm * 1e3}
    for name, pre in (("jacobi", ob.PRECOND_JACOBI), ("two_level", ob.PRECOND_TWO_LEVEL)):
        ctx.set_U(loc(np.zeros(mesh.n_nodes * 3)))
        ctx.newton_step(pre, 1e-8, cg_maxiter=2)                 # warm-up of the solver variant
        ctx.set_U(loc(np.zeros(mesh.n_nodes * 3)))
        info = ctx.newton_step(pre, 1e-8)
        U = ctx.get_U().reshape(-1, 3)[:n_own]
        xyz = mesh.xyz if l2g is None else mesh.xyz[l2g[:n_own]]
        r = np.linalg.norm(xyz[:, :2], axis=1)
        ur = (U[:, :2] * xyz[:, :2] / r[:, None]).sum(axis=1)
        A = (1 + nu) * (1 - 2 * nu) * Ri ** 2 * p / (E * (Re ** 2 - Ri ** 2))
        B = (1 + nu) * Ri ** 2 * Re ** 2 * p / (E * (Re ** 2 - Ri ** 2))
        lame = A * r + B / r
        e_lame, uz, lmax = allmax(np.abs(ur - lame).max(), np.abs(U[:, 2]).max(), lame.max())
        info2 = ctx.step(pre, 1e-8, cg_maxiter=1, update_U=False) if False else None
        ctx.assemble()
        chk = ctx.step(ob.PRECOND_JACOBI, cg_maxiter=1, update_U=False)   # residual of the solved state (linear material: one step solves it)
        out["newton_step_" + name] = {"ms": allmax(info.ms_assemble + info.ms_solve)[0], "cg_iters": int(info.cg_iters),
                                      "us_per_cg_iteration": 1e3 * info.ms_solve / max(int(info.cg_iters), 1),
                                      "radial_error_vs_lame_rel": e_lame / lmax, "uz_rel": uz / lmax,
                                      "residual_after_the_step_rel": chk.norm_r / chk.norm_Fext}
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()


def c5(n=112):
    """examples/clamped_truss / von_misses_truss scaled to a braced cubic lattice of n^3 cells (112 -> 9.95 M bars, Green strain,
    E = 210e9, A = 2.5e-3): (i) under a homogeneous stretch every bar's strain is the closed form and interior nodes are in
    equilibrium; (ii) large-displacement Newton: clamped at x = 0, pulled and sheared at x = L, iterated to a 1e-8 residual."""
    mesh = mg.truss_lattice(n, n, n, 2.0)
    E, A = 210e9, 2.5e-3
    nn = mesh.n_nodes
    fixed = {c: mesh.node_sets["x0"] for c in range(3)}
    free = mg.free_dofs_from_fixed(nn, 3, fixed)
    ctx, l2g, n_own = make(mesh.xyz, [ob.MAT_SVK], [[0.0, E / 2]], free, trusses=mesh.bars, area=np.full(mesh.n_bars, A), strain=ob.STRAIN_GREEN)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    loc = (lambda v: v) if l2g is None else (lambda v: v.reshape(-1, 3)[l2g].ravel())
    eps = 1e-3
    U = np.zeros((nn, 3))
    U[:, 0] = eps * mesh.xyz[:, 0]
    ctx.set_U(loc(U.ravel()))
    ms_asm = timed_assembly(ctx, stream)
    Fint = ctx.get_Fint().reshape(-1, 3)[:n_own]
    xyz = mesh.xyz if l2g is None else mesh.xyz[l2g[:n_own]]
    g = np.rint(xyz / 2.0).astype(int)
    interior = np.all((g > 0) & (g < n), axis=1)
    eq = allmax(np.abs(Fint[interior]).max() / (E * A * eps))[0]
    out = {"config": f"configs[4] braced space-truss lattice {n}^3 cells, Green strain", "n_gpus": world, "n_bars": int(mesh.n_bars), "n_dofs": int(nn * 3),
           "assembly_ms": ms_asm, "bars_per_s": mesh.n_bars / ms_asm * 1e3, "algo_GBps_per_gpu": 456 * mesh.n_bars / world / ms_asm / 1e6,
           "interior_equilibrium_under_homogeneous_stretch": eq}
    F = np.zeros((nn, 3))
    F[mesh.node_sets["x1"], 0] = 0.02 * E * A
    F[mesh.node_sets["x1"], 2] = 0.002 * E * A
    ctx.set_Fext(loc(F.ravel()))
    ctx.set_U(loc(np.zeros(nn * 3)))
    steps, t0 = [], time.perf_counter()
    for it in range(12):
        info = ctx.newton_step(ob.PRECOND_TWO_LEVEL, 1e-10)
        steps.append({"rel_residual_in": info.norm_r / info.norm_Fext, "rel_dU": info.norm_dU / max(info.norm_U, 1e-300), "cg_iters": int(info.cg_iters),
                      "ms": allmax(info.ms_assemble + info.ms_solve)[0]})
        if it > 0 and steps[-1]["rel_residual_in"] < 1e-8:
            break
    out["large_displacement_newton"] = {"load": "x = L face: 0.02 EA per node along x, 0.002 EA along z; x = 0 face clamped", "precond": "two_level",
                                        "cg_reltol": 1e-10, "steps": steps, "wall_s": time.perf_counter() - t0,
                                        "max_displacement_over_L": allmax(np.abs(ctx.get_U()[:n_own * 3]).max())[0] / (2.0 * n)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    {"c3": c3, "c5": c5}[sys.argv[1]]()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
