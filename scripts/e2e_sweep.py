"""End-to-end assembly (host U in, host F_int out) against the number of slice ranges of onsas_assemble_host.
usage: python scripts/e2e_sweep.py [cells]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[(bench.KBULK, bench.MU)], free_dofs=free)
hU, hF = bench.pinned(ctx.n_dofs), bench.pinned(ctx.n_dofs)
hU[:] = U_half
lib, h = ctx._lib, ctx._h
K = 50
for chunks, weight, streams, graph in ((4, 4, 1, 0), (12, 3, 2, 0), (8, 3, 2, 0), (6, 2, 2, 0), (10, 3, 2, 0), (12, 2, 2, 0), (16, 3, 2, 0), (20, 3, 2, 0), (24, 2, 2, 0),
                                       (12, 3, 1, 0), (12, 3, 2, 1), (6, 2, 2, 1), (4, 4, 1, 1), (4, 2, 2, 1), (12, 3, 2, 0), (8, 3, 2, 0), (16, 3, 2, 0)):
    ctx.set_option(ob._lib.OPT_HOST_GRAPH, graph)
    ctx.set_option(ob._lib.OPT_HOST_STREAMS, streams)
    ctx.set_option(ob._lib.OPT_HOST_CHUNKS, chunks)
    ctx.set_option(ob._lib.OPT_HOST_MID_WEIGHT, weight)
    for _ in range(5):
        assert lib.onsas_assemble_host(h, hU, hF) == 0
    t0 = time.perf_counter()
    for _ in range(K):
        assert lib.onsas_assemble_host(h, hU, hF) == 0
    ms = (time.perf_counter() - t0) * 1e3 / K
    print(f"chunks={chunks:3d} mid_weight={weight:2d} streams={streams} graph={graph}  {ms:.4f} ms per call  {mesh.n_tets / ms / 1e6:.3f} G tets/s", flush=True)
t0 = time.perf_counter()
for _ in range(K):
    assert lib.onsas_set_U(h, hU) == 0
    assert lib.onsas_assemble(h) == 0
    assert lib.onsas_get_Fint(h, hF) == 0
ms = (time.perf_counter() - t0) * 1e3 / K
print(f"three calls  {ms:.4f} ms per step  {mesh.n_tets / ms / 1e6:.3f} G tets/s")
# pageable host buffers (no overlap possible): still correct, and how much slower
pU, pF = np.array(hU), np.empty_like(np.asarray(hF))
ctx.set_option(ob._lib.OPT_HOST_GRAPH, 0)
ctx.set_option(ob._lib.OPT_HOST_CHUNKS, 4)
ctx.set_option(ob._lib.OPT_HOST_MID_WEIGHT, 4)
t0 = time.perf_counter()
for _ in range(10):
    assert lib.onsas_assemble_host(h, pU, pF) == 0
print(f"pageable buffers, 4 chunks: {(time.perf_counter() - t0) * 1e2:.4f} ms per call; equal = {np.array_equal(pF, np.asarray(hF))}")
