"""Times one Newton step's linear solve (Jacobi-PCG, 55^3-cell NeoHookean cube) for the persistent CG:
resident CTAs per SM x MB of K pinned in L2, against the legacy persistent kernel, with the per-phase cycle
breakdown of the profiling variant and the single-rank run of the multi-GPU code path (same kernel, run-time switch).

    python scripts/cg_sweep.py [cells] [quick]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
quick = len(sys.argv) > 2
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)


def make(force_mg=0):
    ctx = ob.DeviceContext(0)
    ctx.set_option(L.OPT_FORCE_MG, force_mg)
    ctx.set_nodes(mesh.xyz)
    ctx.set_materials([ob.MAT_NEOHOOKEAN], [[bench.KBULK, bench.MU]])
    ctx.set_tets(mesh.tets)
    ctx.set_free_dofs(free)
    ctx.finalize()
    if force_mg:
        h, off = ctx.p2p_export()
        ctx.p2p_import([h], [off], np.zeros(0, np.int64))
    ctx.set_Fext(Fext)
    return ctx


def run(ctx, tag, mode, bps, prof=0, reps=2):
    ctx.set_option(L.OPT_CG_MODE, mode)
    ctx.set_option(L.OPT_CG_BLOCKS_PER_SM, bps)
    ctx.set_option(L.OPT_CG_PROFILE, prof)
    best = None
    for _ in range(reps):
        ctx.set_U(U_prev)
        info = ctx.newton_step(ob.PRECOND_JACOBI)
        if best is None or info.ms_solve < best.ms_solve:
            best = info
    line = (f"{tag} mode={mode} bps={bps} prof={prof} cg_iters={best.cg_iters} ms_solve={best.ms_solve:8.2f} "
            f"us/iter={1e3 * best.ms_solve / max(best.cg_iters, 1):6.2f} |dU|={best.norm_dU:.12e}")
    if prof:
        pv = ctx.cg_profile()
        slow = pv.pop("slowest_cta_spmv", 0)
        tot = sum(pv.values())
        line += " | " + " ".join(f"{k}={100 * v / tot:.1f}%" for k, v in pv.items()) + f" cyc/iter={tot / max(best.cg_iters, 1):.0f} slowest_cta_spmv/iter={slow / max(best.cg_iters, 1):.0f}"
    print(line, flush=True)


for force in (0, 1):
    ctx = make(force)
    tag = "MG code path (1 rank)" if force else "single GPU           "
    run(ctx, tag + " streamed  ", 0, 4)
    run(ctx, tag + " streamed  ", 0, 4, prof=1, reps=1)
    for bps in (4, 5, 6):
        run(ctx, tag + " registers ", 2, bps)
    run(ctx, tag + " registers ", 2, 6, prof=1, reps=1)
    run(ctx, tag + " multi-launch", 1, 4, reps=1)
    ctx.close()
