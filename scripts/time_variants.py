"""Times the assembly kernel register-budget variants and materials with CUDA events (device-resident)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

minb = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 55
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
lam, G = bench.E_MOD * bench.NU / ((1 + bench.NU) * (1 - 2 * bench.NU)), bench.MU
stream = torch.cuda.Stream()
for mat, (kind, params) in {"neo": (ob.MAT_NEOHOOKEAN, (bench.KBULK, bench.MU)), "svk": (ob.MAT_SVK, (lam, G)),
                            "iso": (ob.MAT_ISOLINEAR, (bench.E_MOD, bench.NU))}.items():
    ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[kind], mat_params=[params], free_dofs=free)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_option(ob._lib.OPT_ASM_MINBLOCKS, minb)
    ctx.set_U(U_half)
    for _ in range(3):
        ctx.assemble()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        ctx.assemble()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"minb={minb} mat={mat} cells={cells} tets={mesh.n_tets} ms={ms:.4f} Gtets/s={mesh.n_tets / ms / 1e6:.3f} "
          f"algoGB/s={1600 * mesh.n_tets / ms / 1e6:.0f}")
    ctx.close()
