"""Diagnostics: the multi-GPU CG kernel run with ONE rank (self-only LL slots, no halo) vs the single-GPU kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
mesh, free, U_half, U_prev, Fext = bench.build_problem(55, 1)
for force in (0, 1):
    ctx = ob.DeviceContext(0)
    ctx.set_option(L.OPT_FORCE_MG, force)
    ctx.set_nodes(mesh.xyz)
    ctx.set_materials([ob.MAT_NEOHOOKEAN], [[bench.KBULK, bench.MU]])
    ctx.set_tets(mesh.tets)
    ctx.set_free_dofs(free)
    ctx.finalize()
    if force:
        h, off = ctx.p2p_export()
        ctx.p2p_import([h], [off], np.zeros(0, np.int64))
    ctx.set_Fext(Fext)
    for bps, prof in ((6, 0), (4, 0), (4, 1)):
        ctx.set_option(L.OPT_CG_BLOCKS_PER_SM, bps)
        ctx.set_option(L.OPT_CG_PROFILE, prof)
        ctx.set_U(U_prev)
        info = ctx.newton_step(ob.PRECOND_JACOBI)
        line = f"force_mg={force} bps={bps} prof={prof} cg_iters={info.cg_iters} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f}"
        if prof:
            pv = ctx.cg_profile()
            slow = pv.pop("slowest_cta_spmv")
            tot = sum(pv.values())
            line += " | " + " ".join(f"{100 * v / tot:.1f}%" for v in pv.values()) + f" slowest_cta_spmv/iter={slow / info.cg_iters:.0f} cycles/iter={tot / info.cg_iters:.0f}"
        print(line, flush=True)
    ctx.close()
