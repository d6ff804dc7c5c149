"""Under torchrun: per-phase cycle breakdown of the peer-memory persistent CG (block 0's view) on the bench problem."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402
from onsas_jl_b200 import multigpu, partition as pt  # noqa: E402

rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
mesh, free, U_half, U_prev, Fext = bench.build_problem(55, world)
order, ranges = pt.rcb_order(mesh.xyz, world)
xyz, tets, inv = pt.renumber(order, mesh.xyz, mesh.tets)
gfree = np.sort(inv[free // 3] * 3 + free % 3)
part = pt.build_local_part(rank, ranges, xyz, tets=tets, free_dofs=gfree)
ctx = multigpu.make_distributed_context(part, [ob.MAT_NEOHOOKEAN], [[bench.KBULK, bench.MU]], dist, local_rank)
loc = lambda v: part.scatter_global(v.reshape(-1, 3)[order].ravel(), 3)  # noqa: E731
ctx.set_Fext(loc(Fext))
L = ob._lib
for bps, prof in ((6, 0), (4, 0), (4, 1)):
    ctx.set_option(L.OPT_CG_BLOCKS_PER_SM, bps)
    ctx.set_option(L.OPT_CG_PROFILE, prof)
    ctx.set_U(loc(U_prev))
    dist.barrier()
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    line = f"rank {rank} bps={bps} prof={prof} cg_iters={info.cg_iters} ms_solve={info.ms_solve:.2f} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f}"
    if prof:
        os.environ["ONSAS_PROF_DUMP"] = f"gpurun_out/mg_cta_rank{rank}.txt"
        pv = ctx.cg_profile()
        os.environ.pop("ONSAS_PROF_DUMP")
        slow = pv.pop("slowest_cta_spmv")
        p = list(pv.values())
        names = ["update_p(+halo)", "sync", "spmv", "sync+allreduce(pAp)", "update_xr+push", "sync+sums", "allreduce(rr,rz)"]
        tot = sum(p)
        line += " | " + " ".join(f"{n}={100 * v / tot:.1f}%" for n, v in zip(names, p)) + f" cycles/iter={tot / info.cg_iters:.0f} slowest_cta_spmv_cycles/iter={slow / info.cg_iters:.0f}"
    print(line, flush=True)
dist.barrier()
# the same local matrix solved by the single-GPU persistent kernel (halo columns present, no communication): isolates
# the cost of the partition ordering from the cost of the multi-GPU machinery
c1 = ob.DeviceContext(local_rank)
c1.set_nodes(part.xyz, part.n_owned)
c1.set_materials([ob.MAT_NEOHOOKEAN], [[bench.KBULK, bench.MU]])
c1.set_tets(part.tets)
c1.set_free_dofs(part.free_dofs, part.n_free_global)
c1.finalize()
c1.set_U(loc(U_prev))
c1.set_Fext(loc(Fext))
for bps, prof in ((6, 0), (4, 1)):
    c1.set_option(L.OPT_CG_BLOCKS_PER_SM, bps)
    c1.set_option(L.OPT_CG_PROFILE, prof)
    c1.assemble()
    info = c1.step(ob.PRECOND_JACOBI, cg_maxiter=1000, update_U=False)
    line = f"rank {rank} LOCAL-ONLY bps={bps} prof={prof} cg_iters={info.cg_iters} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f}"
    if prof:
        pv = c1.cg_profile()
        pv.pop("slowest_cta_spmv")
        tot = sum(pv.values())
        line += " | " + " ".join(f"{k}={100 * v / tot:.1f}%" for k, v in pv.items())
    print(line, flush=True)
dist.barrier()
dist.destroy_process_group()
