"""Unstructured-numbering benchmark: the configs[1] mesh (55^3 cells, NeoHookean) with
  (a) the structured lexicographic numbering every other timing uses,
  (b) node AND element ids randomly permuted (what a Gmsh mesh looks like to the library, Interfaces/Gmsh.jl:25-83),
  (c) the permuted mesh with ONSAS_OPT_REORDER = 1 (Z-curve renumbering inside onsas_finalize_mesh, invisible to the caller),
  (d) the structured mesh with ONSAS_OPT_REORDER = 1.
Reports assembly ms / tets/s, BSELL padding, distinct nodes per slice, us per CG iteration (Jacobi), Newton-step ms.
usage: python scripts/numbering_bench.py [cells]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
rng = np.random.default_rng(2024)
perm, eperm = rng.permutation(mesh.n_nodes), rng.permutation(mesh.n_tets)


def permuted(v):
    out = np.empty_like(v.reshape(-1, 3))
    out[perm] = v.reshape(-1, 3)
    return out.ravel()


cases = {
    "structured": (mesh.xyz, mesh.tets, free, lambda v: v, 0),
    "random": (permuted(mesh.xyz).reshape(-1, 3), perm[mesh.tets[eperm]].astype(np.int32), np.sort(perm[free // 3] * 3 + free % 3), permuted, 0),
    "random+reorder": (permuted(mesh.xyz).reshape(-1, 3), perm[mesh.tets[eperm]].astype(np.int32), np.sort(perm[free // 3] * 3 + free % 3), permuted, 1),
    "structured+reorder": (mesh.xyz, mesh.tets, free, lambda v: v, 1),
    "structured+aggregate-major": (mesh.xyz, mesh.tets, free, lambda v: v, 2),
    "random+aggregate-major": (permuted(mesh.xyz).reshape(-1, 3), perm[mesh.tets[eperm]].astype(np.int32), np.sort(perm[free // 3] * 3 + free % 3), permuted, 2),
}
base = None
for name, (xyz, tets, fr, tr, reorder) in cases.items():
    ctx = ob.context_from_flat(xyz, tets=tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=fr, reorder=reorder)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_U(tr(U_half))
    ctx.set_Fext(tr(Fext))
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
    st = ctx.table_stats()
    ctx.set_U(tr(U_prev))
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    ctx.set_U(tr(U_prev))
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    ctx.set_U(tr(U_prev))
    info2 = ctx.newton_step(ob.PRECOND_TWO_LEVEL)
    ctx.set_U(tr(U_prev))
    info2 = ctx.newton_step(ob.PRECOND_TWO_LEVEL)
    hU = bench.pinned(ctx.n_dofs)
    hF = bench.pinned(ctx.n_dofs)
    hU[:] = tr(U_half)
    import time
    for _ in range(3):
        ctx.assemble_host(hU, hF)
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.assemble_host(hU, hF)
    ms_e2e = (time.perf_counter() - t0) * 1e3 / 20
    rec = {"numbering": name, "assembly_ms": ms, "tets_per_s": mesh.n_tets / (ms * 1e-3), "bsell_padding": st["padded_block_slots"] / st["nnz_blocks"] - 1.0,
           "max_pairs_per_slice": st["max_pairs_per_slice"], "newton_step_ms": info.ms_assemble + info.ms_solve, "cg_iters": int(info.cg_iters),
           "us_per_cg_iteration": 1e3 * info.ms_solve / max(int(info.cg_iters), 1), "e2e_ms": ms_e2e,
           "two_level_newton_step_ms": info2.ms_assemble + info2.ms_solve, "two_level_cg_iters": int(info2.cg_iters),
           "two_level_us_per_cg_iteration": 1e3 * info2.ms_solve / max(int(info2.cg_iters), 1)}
    if base is None:
        base = rec
    rec["assembly_vs_structured"] = base["assembly_ms"] / ms
    rec["cg_iteration_vs_structured"] = base["us_per_cg_iteration"] / rec["us_per_cg_iteration"]
    print(json.dumps(rec), flush=True)
    ctx.close()
