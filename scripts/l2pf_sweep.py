"""Sweep of ONSAS_OPT_CG_L2_PREFETCH (slices per consumer warp the producer warp of cg_stream pulls into L2 behind the ring
between SpMV phases) on configs[1]: Jacobi (single-reduction) and two-level PCG; the solution must not change by a bit.
usage: python scripts/l2pf_sweep.py [cells] [pf values ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
pfs = [int(a) for a in sys.argv[2:]] or [0, 2, 3, 4, 5, 6, 8, 10, 0]
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
ctx.set_Fext(Fext)
for pre, name in ((ob.PRECOND_JACOBI, "jacobi"), (ob.PRECOND_TWO_LEVEL, "two_level")):
    ref = None
    for pf in pfs:
        ctx.set_option(L.OPT_CG_L2_PREFETCH, pf)
        best = None
        for rep in range(3):
            ctx.set_option(L.OPT_CG_PROFILE, 1 if rep == 2 else 0)
            ctx.set_U(U_prev)
            info = ctx.newton_step(pre)
            if rep < 2:
                best = info.ms_solve if best is None else min(best, info.ms_solve)
        U = ctx.get_U()
        if ref is None:
            ref = U
        pv = ctx.cg_profile()
        pv.pop("slowest_cta_spmv", 0)
        print(f"{name} cells={cells} l2_prefetch={pf} cg_iters={info.cg_iters} ms_solve={best:.2f} us/iter={1e3 * best / info.cg_iters:.2f} "
              f"bitwise_same={bool(np.array_equal(U, ref))} sha1(U)={__import__('hashlib').sha1(U.tobytes()).hexdigest()[:12]} | " + " ".join(f"{k}={v / info.cg_iters:.0f}" for k, v in pv.items()), flush=True)
    ctx.set_option(L.OPT_CG_PROFILE, 0)
