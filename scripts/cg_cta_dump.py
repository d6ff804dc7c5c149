"""Diagnostics: per-CTA SpMV cycles (and SM ids) of one profiled fused-CG solve -> gpurun_out/<tag>/cta_<name>.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
out = sys.argv[1]
mesh, free, U_half, U_prev, Fext = bench.build_problem(55, 1)
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
ctx.set_Fext(Fext)
for bps in (4, 6):
    ctx.set_option(L.OPT_CG_BLOCKS_PER_SM, bps)
    ctx.set_option(L.OPT_CG_PROFILE, 1)
    ctx.set_U(U_prev)
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    os.environ["ONSAS_PROF_DUMP"] = os.path.join(out, f"cta_fused_bps{bps}.txt")
    pv = ctx.cg_profile()
    print(bps, info.cg_iters, info.ms_solve, pv, flush=True)
