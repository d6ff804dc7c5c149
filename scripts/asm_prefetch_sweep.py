"""Assembly kernel: L2 prefetch distance sweep (ONSAS_OPT_ASM_PREFETCH), CUDA events, device-resident state.
usage: python scripts/asm_prefetch_sweep.py [cells]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
stream = torch.cuda.Stream()
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[(bench.KBULK, bench.MU)], free_dofs=free)
ctx.set_stream(stream.cuda_stream)
ctx.set_U(U_half)
ref = None
for rep in range(2):
    for dist in (0, 148, 296, 444, 518, 592, 740, 888, 1332, 2220, 0):
        ctx.set_option(ob._lib.OPT_ASM_PREFETCH, dist)
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
        same = ""
        if rep == 0:
            F, K = ctx.get_Fint(), ctx.get_csr()[2]
            if ref is None:
                ref = (F, K)
            same = f" bitwise_equal_to_dist0={np.array_equal(F, ref[0]) and np.array_equal(K, ref[1])}"
        print(f"prefetch_dist={dist:5d} ms={ms:.4f} Gtets/s={mesh.n_tets / ms / 1e6:.3f}{same}", flush=True)
