"""Per-warp SpMV cycles of the streamed CG (ONSAS_PROF_DUMP): where does the barrier skew behind the SpMV phase come from --
the 13-against-12 slices of the static dealing or SM-dependent HBM service?  usage: python scripts/spmv_skew.py [cells]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 55
dump = "/tmp/onsas_prof_dump.txt"
os.environ["ONSAS_PROF_DUMP"] = dump
mesh, free, U_half, U_prev, Fext = bench.build_problem(cells, 1)
ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=[ob.MAT_NEOHOOKEAN], mat_params=[[bench.KBULK, bench.MU]], free_dofs=free)
ctx.set_Fext(Fext)
ctx.set_option(L.OPT_CG_PROFILE, 1)
for rep in range(2):
    ctx.set_U(U_prev)
    info = ctx.newton_step(ob.PRECOND_JACOBI)
    pv = ctx.cg_profile()
ns = ctx.table_stats()["n_slices"]
v = np.loadtxt(dump)[: 148 * 12].reshape(148, 12) / (info.cg_iters + 1)
first = np.arange(148)[:, None] * 12 + np.arange(12)[None, :]
n_my = (ns - 1 - first) // (148 * 12) + 1
print(f"cg_iters={info.cg_iters} us/iter={1e3 * info.ms_solve / info.cg_iters:.2f}", {k: round(x / info.cg_iters) for k, x in pv.items()})
print(f"slices={ns}; warps with {n_my.max()} slices: {(n_my == n_my.max()).sum()}, with {n_my.min()}: {(n_my == n_my.min()).sum()}")
for m in np.unique(n_my):
    w = v[n_my == m]
    print(f"  warps with {m} slices: cycles per SpMV phase mean {w.mean():.0f} min {w.min():.0f} max {w.max():.0f} std {w.std():.0f}")
cta = v.max(axis=1)
print(f"per-CTA slowest warp: mean {cta.mean():.0f} min {cta.min():.0f} max {cta.max():.0f}; sorted deciles", np.percentile(cta, [0, 10, 25, 50, 75, 90, 100]).round())
print("per-CTA slowest warp by CTA index (x100 cycles):", " ".join(f"{x / 100:.0f}" for x in cta))
