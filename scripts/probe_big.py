"""Per-phase cycles of the streamed CG on a per-rank-sized piece of configs[3] (default 94^3 cells = 4.98 M SVK tets = what one
of 8 GPUs holds of the 39.9 M-tet cube): Jacobi vs two-level, fused vs stand-alone w pass, coarse set-up cost.
usage: python scripts/probe_big.py [cells]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ONSAS_VERBOSE"] = "1"
os.environ["ONSAS_PROF_VERBOSE"] = "1"
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402
from onsas_jl_b200 import meshgen as mg  # noqa: E402

L = ob._lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 94
reorder = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mesh = mg.box_tet_mesh(n, n, n, 1.0, 1.0, 1.0)
free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
a0, b0 = bench.uniaxial_state(bench._P_svk, 3.0 * 7 / 8, (1.8, 0.5))
U_prev = mg.homogeneous_field(mesh.xyz, a0, b0)
Fext = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (3.0, 0.0, 0.0))
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_SVK], mat_params=[[bench.LAM, bench.MU]], free_dofs=free, reorder=reorder)
ctx.set_Fext(Fext)
print("reorder", reorder, "tets", mesh.n_tets, "tables", ctx.table_stats(), flush=True)
for precond, fused in ((ob.PRECOND_JACOBI, 0), (ob.PRECOND_TWO_LEVEL, 1), (ob.PRECOND_TWO_LEVEL, 0)):
    ctx.set_option(L.OPT_COARSE_FUSED, fused)
    for prof in (0, 1):
        ctx.set_option(L.OPT_CG_PROFILE, prof)
        ctx.set_U(U_prev)
        info = ctx.newton_step(precond)
        line = (f"precond={precond} fused={fused} prof={prof} cg_iters={info.cg_iters} ms_assemble={info.ms_assemble:.3f} ms_solve={info.ms_solve:.2f} "
                f"us/iter={1e3 * info.ms_solve / info.cg_iters:.2f}")
        if prof:
            pv = ctx.cg_profile()
            slow = pv.pop("slowest_cta_spmv", 0)
            tot = sum(pv.values())
            line += " | " + " ".join(f"{k}={v / info.cg_iters:.0f}" for k, v in pv.items()) + f" cyc/iter={tot / info.cg_iters:.0f}"
        print(line, flush=True)
b = np.random.default_rng(0).standard_normal(mesh.n_nodes * 3)
ctx.set_option(L.OPT_CG_PROFILE, 0)
ctx.set_option(L.OPT_COARSE_FUSED, 1)
ctx.set_U(U_prev)
ctx.assemble()
ctx.synchronize()
for k in range(3):
    t0 = time.perf_counter()
    x, its, res = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-3)
    print(f"two-level pcg call {k} (the first one includes the coarse set-up): {1e3 * (time.perf_counter() - t0):.2f} ms, {its} iterations", flush=True)
