"""Diagnostics of the streamed CG: per-phase cycles for ring
geometries (ONSAS_STREAM_CW / ONSAS_STREAM_DEPTH experiment knobs)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    os.environ["ONSAS_VERBOSE"] = "1"
    import bench  # noqa: E402
    import onsas_jl_b200 as ob  # noqa: E402

    L = ob._lib
    mesh, free, U_half, U_prev, Fext = bench.build_problem(55, 1)
    ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
    ctx.set_Fext(Fext)
    for precond, rbm, fused in ((ob.PRECOND_JACOBI, 0, 0), (ob.PRECOND_TWO_LEVEL, 0, 0), (ob.PRECOND_TWO_LEVEL, 0, 1),
                                (ob.PRECOND_TWO_LEVEL, 1, 1), (ob.PRECOND_TWO_LEVEL, 1, 0)):
        ctx.set_option(L.OPT_COARSE_RBM, rbm)
        ctx.set_option(L.OPT_COARSE_FUSED, fused)
        for prof in (0, 0, 1):
            ctx.set_option(L.OPT_CG_PROFILE, prof)
            ctx.set_U(U_prev)
            info = ctx.newton_step(precond)
            line = (f"precond={precond} rbm={rbm} fused={fused} prof={prof} "
                    f"cg_iters={info.cg_iters} ms_solve={info.ms_solve:.2f} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f} |dU|={info.norm_dU:.10e}")
            if prof:
                pv = ctx.cg_profile()
                slow = pv.pop("slowest_cta_spmv", 0)
                tot = sum(pv.values())
                line += " | " + " ".join(f"{k}={v / info.cg_iters:.0f}" for k, v in pv.items()) + f" cyc/iter={tot / info.cg_iters:.0f} slowest_warp_spmv/iter={slow / info.cg_iters:.0f}"
            print(line, flush=True)
    # set-up cost of the two-level preconditioner: the first solve after an assembly builds and inverts E
    import time
    import numpy as np
    b = np.random.default_rng(0).standard_normal(mesh.n_nodes * 3)
    ctx.set_option(L.OPT_CG_PROFILE, 0)
    ctx.set_U(U_prev)
    ctx.assemble()
    ctx.synchronize()
    for k in range(3):
        t0 = time.perf_counter()
        x, its, res = ctx.pcg(b, ob.PRECOND_TWO_LEVEL, 1e-6)
        print(f"two-level pcg call {k} (first one includes the coarse set-up): {1e3 * (time.perf_counter() - t0):.2f} ms, {its} iterations", flush=True)
else:
    for cw, depth in ((12, 2),):
        env = dict(os.environ, ONSAS_STREAM_CW=str(cw), ONSAS_STREAM_DEPTH=str(depth))
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, timeout=300)
