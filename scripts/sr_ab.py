"""A/B of the CG recurrences of the streamed persistent solver (ONSAS_OPT_CG_SINGLE_REDUCTION) on configs[1]:
classic (3 grid barriers, 2 reductions per iteration) vs single-reduction (2 barriers, 1 reduction), Jacobi preconditioner.
usage: python scripts/sr_ab.py [cells]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
ctx.set_Fext(Fext)
for pre, name in ((ob.PRECOND_JACOBI, "jacobi"), (ob.PRECOND_TWO_LEVEL, "two_level")):
    ref = None
    for sr in (0, 3, 0, 3):
        ctx.set_option(L.OPT_CG_SINGLE_REDUCTION, sr)
        for prof in (0, 1):
            ctx.set_option(L.OPT_CG_PROFILE, prof)
            ctx.set_U(U_prev)
            info = ctx.newton_step(pre)
            U = ctx.get_U()
            if ref is None:
                ref = U
            line = (f"{name} single_reduction={sr} prof={prof} cg_iters={info.cg_iters} ms_solve={info.ms_solve:.2f} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f} "
                    f"|dU|={info.norm_dU:.12e} max|U-U_classic|={np.abs(U - ref).max():.2e}")
            if prof:
                pv = ctx.cg_profile()
                pv.pop("slowest_cta_spmv", 0)
                line += " | " + " ".join(f"{k}={v / info.cg_iters:.0f}" for k, v in pv.items())
            print(line, flush=True)
