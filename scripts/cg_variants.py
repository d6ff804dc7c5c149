"""Times one Newton step (assemble + Jacobi-PCG) for the persistent-CG occupancy variants, with the per-phase
cycle breakdown of the profiling variant, and the multi-launch driver for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
L = ob._lib
for bps in (4, 5, 6):
    ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
    ctx.set_option(L.OPT_CG_BLOCKS_PER_SM, bps)
    ctx.set_Fext(Fext)
    for prof in (0, 0, 1):
        ctx.set_option(L.OPT_CG_PROFILE, prof)
        ctx.set_U(U_prev)
        info = ctx.newton_step(ob.PRECOND_JACOBI)
        line = f"bps={bps} prof={prof} grid={ctx.table_stats()['cg_grid']} cg_iters={info.cg_iters} ms_solve={info.ms_solve:.2f} us/iter={1e3 * info.ms_solve / max(info.cg_iters, 1):.2f}"
        if prof:
            p = ctx.cg_profile()
            tot = sum(p.values())
            line += " | " + " ".join(f"{k}={100 * v / tot:.1f}%" for k, v in p.items()) + f" cycles/iter={tot / max(info.cg_iters, 1):.0f}"
        print(line, flush=True)
    ctx.close()
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
ctx.set_option(L.OPT_CG_MODE, 1)
ctx.set_Fext(Fext)
for ce in (16, 64):
    ctx.set_option(L.OPT_CG_CHECK_EVERY, ce)
    ctx.set_U(U_prev)
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    print(f"multi-launch check_every={ce} cg_iters={info.cg_iters} ms_solve={info.ms_solve:.2f} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f}", flush=True)
