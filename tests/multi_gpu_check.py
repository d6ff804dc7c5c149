"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
Every rank builds the same small global problem, keeps its RCB part, and the distributed assembly / SpMV / PCG /
Newton solve are compared with the oracle on the global mesh (F_int, K rows, U within the north-star tolerances)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import onsas_jl_b200 as ob  # noqa: E402
from onsas_jl_b200 import meshgen as mg  # noqa: E402
from onsas_jl_b200 import partition as pt  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import cases  # noqa: E402


def make_ctx(part, kind, params, rank, world, local_rank):
    from onsas_jl_b200 import multigpu
    return multigpu.make_distributed_context(part, kind, params, dist, local_rank, p2p=True)


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    m, mesh = cases.box_model(12, 6, 6, mat="neo", jitter=0.1)
    order, ranges = pt.rcb_order(m.xyz, world)
    xyz, tets, inv = pt.renumber(order, m.xyz, m.tets)
    free = np.sort(inv[m.free_dofs // 3] * 3 + m.free_dofs % 3)
    gm = O.FlatModel(xyz=xyz, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    part = pt.build_local_part(rank, ranges, xyz, tets=tets, free_dofs=free)
    ctx = make_ctx(part, m.mat_kind, m.mat_params, rank, world, local_rank)
    own = part.owned_global_dofs(3)

    # ---- assembly: owned rows of F_int and K equal the global oracle rows
    U = cases.random_U(gm, 0.02)
    ref = O.Assembly(gm).assemble(U)
    ctx.set_U(part.scatter_global(U, 3))
    ctx.assemble()
    Fl = ctx.get_Fint()[:len(own)]
    e_f = cases.rel_err(Fl, ref.F_int[own])
    rp, ci, v = ctx.get_csr()
    import scipy.sparse as sp
    Kl = sp.csr_matrix((v, ci, rp), shape=(len(own), part.n_local * 3)).tocoo()
    gcols = part.local_to_global[Kl.col // 3] * 3 + Kl.col % 3
    Kg = sp.csr_matrix((Kl.data, (Kl.row, gcols)), shape=(len(own), gm.n_dofs))
    e_k = abs(Kg - ref.csr()[own]).max() / abs(ref.csr()).max()

    # ---- distributed SpMV and PCG against the oracle
    mask = gm.free_mask()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(gm.n_dofs) * mask
    yl = ctx.spmv(part.scatter_global(x, 3))[:len(own)]
    e_y = cases.rel_err(yl, ((ref.csr() @ x) * mask)[own])
    b = rng.standard_normal(gm.n_dofs)
    d = np.where(mask, ref.csr().diagonal(), 1.0)
    xo, ito, _ = O.cg(ref.rowptr, ref.col, ref.val, mask, b, diag=d, reltol=1e-12)
    e_x, its_modes = 0.0, []
    for mode in (0, 1, 2):   # 0 = TMA-streamed persistent kernel over NVLink peer memory, 1 = one launch per phase + NCCL, 2 = register-fed persistent kernel
        ctx.set_option(ob._lib.OPT_CG_MODE, mode)
        xs, its, res = ctx.pcg(part.scatter_global(b, 3), ob.PRECOND_JACOBI, 1e-12)
        e_x = max(e_x, np.abs(xs[:len(own)] - xo[own]).max() / np.abs(xo).max())
        its_modes.append(its)
    ctx.set_option(ob._lib.OPT_CG_MODE, 0)
    its = its_modes
    # two-level preconditioner on N ranks: every rank's coarse space covers its owned nodes (block-diagonal E)
    e_x2, its2 = float("inf"), -1
    try:
        xs2, its2, _ = ctx.pcg(part.scatter_global(b, 3), ob.PRECOND_TWO_LEVEL, 1e-12)
        e_x2 = np.abs(xs2[:len(own)] - xo[own]).max() / np.abs(xo).max()
    except ob.OnsasError as exc:
        print(f"rank {rank}: two-level PCG failed: {exc}", flush=True)

    # ---- distributed Newton solve of the compression example (9 load steps, tol 1e-10, as the reference ships it)
    #      vs the oracle's direct-solve Newton, on the undistorted mesh of the same grid
    m2, mesh2 = cases.box_model(12, 6, 6, mat="neo")
    order, ranges = pt.rcb_order(m2.xyz, world)          # the partition of THIS mesh (the jittered one above splits elsewhere)
    xyz2, tets, inv = pt.renumber(order, m2.xyz, m2.tets)
    free = np.sort(inv[m2.free_dofs // 3] * 3 + m2.free_dofs % 3)
    gm = O.FlatModel(xyz=xyz2, tets=tets, mat_kind=m.mat_kind, mat_params=m.mat_params, free_dofs=free)
    part = pt.build_local_part(rank, ranges, xyz2, tets=tets, free_dofs=free)
    own = part.owned_global_dofs(3)
    ctx.close()
    ctx = make_ctx(part, m.mat_kind, m.mat_params, rank, world, local_rank)
    Fg = mg.global_face_load(mesh2.n_nodes, mesh2.xyz, mesh2.faces["x1"], (-1.0, 0.0, 0.0)).reshape(-1, 3)[order].ravel()
    tols = O.ConvergenceSettings(1e-10, 1e-10, 20)
    lfs = np.linspace(1 / 9, 1.0, 9)
    refn = O.newton_solve(gm, lfs, lambda t: Fg * t, tols)
    ctx.set_U(np.zeros(part.n_local * 3))
    iters = []
    import math
    for t in lfs:
        ctx.set_Fext(part.scatter_global(Fg * t, 3))
        dU_rel = dr_rel = 1e12
        it = 0
        while O.criterion(dU_rel, dr_rel, it, tols) == "NotConvergedYet":
            info = ctx.newton_step(ob.PRECOND_JACOBI, 1e-13)
            dU_rel = info.norm_dU / info.norm_U if info.norm_U > 0 else math.inf
            dr_rel = info.norm_r / info.norm_Fext
            it += 1
        iters.append(it)
    Uloc = ctx.get_U()
    Ul = Uloc[:len(own)]
    e_u = np.abs(Ul - refn.U[-1][own]).max() / np.abs(refn.U[-1]).max()
    # the solver keeps the halo part of U current without any exchange: it must equal the owners' values bit for bit
    Ug = torch.zeros(gm.n_dofs, dtype=torch.float64, device="cuda")
    Ug[torch.as_tensor(own, device="cuda")] = torch.as_tensor(Ul, device="cuda")
    dist.all_reduce(Ug)
    Ug = Ug.cpu().numpy()
    halo_ok = np.array_equal(Uloc, part.scatter_global(Ug, 3))
    # ---- the same solve through the library's own partitioner (onsas_part_*): caller numbering in, bitwise the same U
    ctx.close()
    P = ob.NativePartition(m2.xyz, world, tets=m2.tets, free_dofs=m2.free_dofs)
    from onsas_jl_b200 import multigpu
    ctx = multigpu.make_distributed_context(P, m.mat_kind, m.mat_params, dist, local_rank, p2p=True)
    l2g = P.local_to_global(rank)
    Fo = mg.global_face_load(mesh2.n_nodes, mesh2.xyz, mesh2.faces["x1"], (-1.0, 0.0, 0.0)).reshape(-1, 3)
    ctx.set_U(np.zeros(len(l2g) * 3))
    iters_n = []
    for t in lfs:
        ctx.set_Fext((Fo[l2g] * t).ravel())
        dU_rel = dr_rel = 1e12
        it = 0
        while O.criterion(dU_rel, dr_rel, it, tols) == "NotConvergedYet":
            info = ctx.newton_step(ob.PRECOND_JACOBI, 1e-13)
            dU_rel = info.norm_dU / info.norm_U if info.norm_U > 0 else math.inf
            dr_rel = info.norm_r / info.norm_Fext
            it += 1
        iters_n.append(it)
    n_own = P.sizes(rank)["n_owned"]
    Un = torch.zeros((mesh2.n_nodes, 3), dtype=torch.float64, device="cuda")
    Un[torch.as_tensor(l2g[:n_own].astype(np.int64), device="cuda")] = torch.as_tensor(ctx.get_U().reshape(-1, 3)[:n_own], device="cuda")
    dist.all_reduce(Un)
    Un = Un.cpu().numpy()
    native_ok = np.array_equal(Un[order].ravel(), Ug) and iters_n == iters
    # the two-level preconditioner with the GLOBAL coarse level (level-2 aggregates across ranks) and without it: both solve
    # the same system to the oracle's answer; the global level must not need more iterations
    rng2 = np.random.default_rng(9)
    b2 = rng2.standard_normal(mesh2.n_nodes * 3)
    ctx.set_U((0.02 * rng2.standard_normal((mesh2.n_nodes, 3)))[l2g].ravel())
    ctx.assemble()
    its_g = {}
    xg = {}
    for glob in (1, 0, 1):
        ctx.set_option(ob._lib.OPT_COARSE_GLOBAL, glob)
        xs, its_, _ = ctx.pcg(b2.reshape(-1, 3)[l2g].ravel(), ob.PRECOND_TWO_LEVEL, 1e-12)
        its_g[glob] = its_
        xg[glob] = xs
    xj, its_j, _ = ctx.pcg(b2.reshape(-1, 3)[l2g].ravel(), ob.PRECOND_JACOBI, 1e-12)
    n_own = P.sizes(rank)["n_owned"]
    sc = max(np.abs(xj).max(), 1e-300)
    e_glob = max(np.abs(xg[1][:3 * n_own] - xj[:3 * n_own]).max(), np.abs(xg[0][:3 * n_own] - xj[:3 * n_own]).max()) / sc
    glob_ok = e_glob < 1e-8 and its_g[1] <= its_g[0] + 5
    if rank == 0:
        print(f"multi-gpu check world={world}: two-level with the global coarse level {its_g[1]} iterations, without {its_g[0]}, jacobi {its_j}; x vs jacobi solve {e_glob:.2e}", flush=True)
    if rank == 0 and os.environ.get("ONSAS_MULTI_DUMP"):
        np.save(os.environ["ONSAS_MULTI_DUMP"], Un)      # caller numbering: compared with the one-process multi-device context
    errs = torch.tensor([e_f, e_k, e_y, e_x, e_u, e_x2, 0.0 if halo_ok else 1.0, 0.0 if native_ok else 1.0, 0.0 if glob_ok else 1.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        e_f, e_k, e_y, e_x, e_u, e_x2, bad_halo, bad_native, bad_glob = errs.tolist()
        halo_ok, native_ok, glob_ok = bad_halo == 0.0, bad_native == 0.0, bad_glob == 0.0
        print(f"multi-gpu check world={world}: two-level pcg x {e_x2:.2e} (its {its2} vs jacobi {its[0]})", flush=True)
        print(f"multi-gpu check world={world}: F_int {e_f:.2e}  K {e_k:.2e}  spmv {e_y:.2e}  pcg x {e_x:.2e} (its {its} vs {ito})  "
              f"newton U {e_u:.2e} iters {iters} vs {refn.iterations}", flush=True)
        print(f"multi-gpu check world={world}: halo part of U bitwise current after the solves: {halo_ok}; native partitioner bitwise equal: {native_ok}", flush=True)
        ok = e_f < 1e-12 and e_k < 1e-12 and e_y < 1e-12 and e_x < 1e-8 and e_u < 1e-8 and iters == refn.iterations and e_x2 < 1e-8
        ok = ok and halo_ok and native_ok and glob_ok
        print(f"multi-gpu check world={world}: global coarse level ok: {glob_ok}", flush=True)
        print("MULTI_GPU_CHECK_OK" if ok else "MULTI_GPU_CHECK_FAILED", flush=True)
    dist.barrier()          # never leave the other ranks waiting on a failed assertion
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
