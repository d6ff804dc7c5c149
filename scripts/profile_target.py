"""Small driver for ncu / compute-sanitizer: a few launches of each hot kernel, no CPU baseline.
usage: python scripts/profile_target.py [cells] [mat] [n_asm] [n_spmv] [newton 0/1] [precond 0/1/2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mat = sys.argv[2] if len(sys.argv) > 2 else "neo"
n_asm = int(sys.argv[3]) if len(sys.argv) > 3 else 4
n_spmv = int(sys.argv[4]) if len(sys.argv) > 4 else 4
newton = int(sys.argv[5]) if len(sys.argv) > 5 else 1
precond = int(sys.argv[6]) if len(sys.argv) > 6 else 1
minb = int(os.environ.get("ONSAS_ASM_MINB", "2"))

mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
lam, G = bench.E_MOD * bench.NU / ((1 + bench.NU) * (1 - 2 * bench.NU)), bench.MU
kind, params = {"neo": (ob.MAT_NEOHOOKEAN, (bench.KBULK, bench.MU)), "svk": (ob.MAT_SVK, (lam, G)),
                "iso": (ob.MAT_ISOLINEAR, (bench.E_MOD, bench.NU))}[mat]
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[kind], mat_params=[params], free_dofs=free)
ctx.set_option(ob._lib.OPT_ASM_MINBLOCKS, minb)
ctx.set_U(U_half)
ctx.set_Fext(Fext)
for _ in range(n_asm):
    ctx.assemble()
ctx.synchronize()
for _ in range(n_spmv):
    ctx.spmv_resident()
ctx.synchronize()
if newton:
    ctx.set_U(U_prev)
    info = ctx.newton_step(precond)
    print("newton step: precond", precond, "cg_iters", info.cg_iters, "ms_assemble", info.ms_assemble, "ms_solve", info.ms_solve)
if newton and os.environ.get("ONSAS_PROFILE_MULTILAUNCH"):
    ctx.set_option(ob._lib.OPT_CG_MODE, 1)
    ctx.set_U(U_prev)
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    print("newton step (multi-launch): cg_iters", info.cg_iters, "ms_assemble", info.ms_assemble, "ms_solve", info.ms_solve)
print("tables", ctx.table_stats())
