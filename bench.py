#!/usr/bin/env python
"""bench.py -- headline benchmark of the ONSAS.jl Newton-Raphson hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells C]

Metric (BASELINE.json): tet f_int + K_t ASSEMBLED elements/s (one "step" = one assemble! pass over the whole
mesh: element evaluation + deterministic assembly of K, F_int, stress, strain) with the Newton-step time
(assemble + Jacobi-PCG + update) reported beside it.
Workload at N = 1 (configs[1]): examples/uniaxial_compression -- NeoHookean tet cube, synthetic structured mesh
of 55^3 cells = 998 250 tetrahedra, evaluated at the analytic homogeneous state of load factor 0.5.
N > 1: weak scaling -- a box of N x 55^3 cells partitioned by recursive coordinate bisection into N slabs,
one rank per GPU (the library's own partitioner, onsas_part_*); value = all tets / max-over-ranks time.
Beside it, at EVERY N, the named multi-GPU configuration (configs[3]) as a STRONG-scaling leg ("strong_c4"): the synthetic
structured SVK cube of 188^3 cells = 39 868 032 tetrahedra partitioned over the N GPUs -- assembly tets/s, Newton-step
time with its CG iterations (Jacobi and two-level) -- and the line carries its own parity: the residual at the analytic
equilibrium state and the distance of the state after the timed Newton step from its analytic value (the problem is
homogeneous, so one exact Newton step from a homogeneous state lands on the homogeneous state given by the Newton
iterate of the 2 x 2 scalar system P11(alpha, beta) = p, P22(alpha, beta) = 0).

Prints ONE JSON line (rank 0).  Timing: CUDA events on the stream the kernels run on, W >= 3 warm-up steps,
inputs larger than L2 (K values + element records + tables ~ 0.4 GB per pass vs 126 MB L2).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# torchrun gives its workers OMP_NUM_THREADS = 1 unless the caller set it: the library's host-side table builder and partitioner
# (OpenMP, once per mesh, outside every timed region) would run serially -- 20 s instead of 5 for the configs[3] mesh.  Share the
# host cores among the ranks of this node instead; must happen before anything loads the OpenMP runtime.
if os.environ.get("LOCAL_WORLD_SIZE") and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // max(1, int(os.environ["LOCAL_WORLD_SIZE"]))))

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_TET = 1600  # 16 conn + 96 X + 96 U + 1152 K_e + 96 f_e + 72 P + 72 C
E_MOD, NU = 1.0, 0.3
MU, KBULK = E_MOD / (2 * (1 + NU)), E_MOD / (3 * (1 - 2 * NU))


def _neo_state(load: float):
    """alpha, beta with P11 = -load, P22 = 0 for the compressible NeoHookean of NeoHookeanMaterial.jl:94-102."""
    a, b = 1.0, 1.0
    for _ in range(60):
        f = np.array([MU * a - MU / a + KBULK * b ** 2 * (a * b ** 2 - 1) + load,
                      MU * b - MU / b + KBULK * b * (a ** 2 * b ** 2 - a)])
        J = np.array([[MU + MU / a ** 2 + KBULK * b ** 4, KBULK * (4 * a * b ** 3 - 2 * b)],
                      [KBULK * b * (2 * a * b ** 2 - 1), MU + MU / b ** 2 + KBULK * (3 * a ** 2 * b ** 2 - a)]])
        d = np.linalg.solve(J, -f)
        a, b = a + d[0], b + d[1]
        if np.abs(d).max() < 1e-15:
            break
    return float(a), float(b)


def _P_neo(a, b):
    """(P11, P22) and their Jacobian for F = diag(a, b, b), compressible NeoHookean (NeoHookeanMaterial.jl:94-102)."""
    f = np.array([MU * a - MU / a + KBULK * b ** 2 * (a * b ** 2 - 1), MU * b - MU / b + KBULK * b * (a ** 2 * b ** 2 - a)])
    J = np.array([[MU + MU / a ** 2 + KBULK * b ** 4, KBULK * (4 * a * b ** 3 - 2 * b)],
                  [KBULK * b * (2 * a * b ** 2 - 1), MU + MU / b ** 2 + KBULK * (3 * a ** 2 * b ** 2 - a)]])
    return f, J


LAM = E_MOD * NU / ((1 + NU) * (1 - 2 * NU))


def _P_svk(a, b):
    """(P11, P22) and their Jacobian for F = diag(a, b, b), SVK: S = lambda tr(E) I + 2 G E (SVKMaterial.jl:89-100), P = F S."""
    E11, E22 = 0.5 * (a * a - 1), 0.5 * (b * b - 1)
    tr = E11 + 2 * E22
    S11, S22 = LAM * tr + 2 * MU * E11, LAM * tr + 2 * MU * E22
    f = np.array([a * S11, b * S22])
    J = np.array([[S11 + a * (LAM + 2 * MU) * a, a * 2 * LAM * b],
                  [b * LAM * a, S22 + b * (2 * LAM + 2 * MU) * b]])
    return f, J


def uniaxial_state(P, traction, start=(1.0, 1.0)):
    """Equilibrium stretches (alpha, beta): P11 = traction, P22 = 0."""
    a, b = start
    for _ in range(100):
        f, J = P(a, b)
        d = np.linalg.solve(J, -(f - np.array([traction, 0.0])))
        a, b = a + d[0], b + d[1]
        if np.abs(d).max() < 1e-15:
            break
    return float(a), float(b)


def newton_iterate(P, traction, a0, b0):
    """The state after ONE exact Newton step of the finite-element problem from the homogeneous state (a0, b0) under the
    uniform traction: the linearised problem of a homogeneous body is solved by a homogeneous increment, which is the
    Newton step of the scalar system."""
    f, J = P(a0, b0)
    d = np.linalg.solve(J, -(f - np.array([traction, 0.0])))
    return float(a0 + d[0]), float(b0 + d[1])


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(cells: int, n_ranks: int):
    """Global mesh (box of n_ranks x cells^3 hexes, unit cells), BCs of the uniaxial example, states."""
    from onsas_jl_b200 import meshgen as mg
    mesh = mg.box_tet_mesh(cells * n_ranks, cells, cells, float(n_ranks), 1.0, 1.0)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    a0, b0 = _neo_state(4.0 / 9.0)
    a1, b1 = _neo_state(0.5)
    U_half = mg.homogeneous_field(mesh.xyz, a1, b1)        # state the assembly is timed at (F != I everywhere)
    U_prev = mg.homogeneous_field(mesh.xyz, a0, b0)        # start of the timed Newton step (previous load step)
    Fext = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (-0.5, 0.0, 0.0))
    return mesh, free, U_half, U_prev, Fext


def workload_config(cells, n_gpus, n_tets, n_dofs):
    """`config` of the JSON line: the workload only (identical in both arms at N = 1), nothing about the implementation."""
    return {"workload": f"examples/uniaxial_compression NeoHookean tet cube (configs[1]), structured {cells}^3 cells per GPU, "
                        "state = analytic homogeneous field at load factor 0.5",
            "n_tets": int(n_tets), "n_dofs": int(n_dofs), "l2": "inputs larger than L2 (~0.4 GB touched per pass)"}


def pinned(n):
    import torch
    return torch.empty(n, dtype=torch.float64).pin_memory().numpy()


def make_context(ob, mesh, free, kind, params, N, rank, dist, local_rank, no_p2p=False):
    """One rank's device context of the global mesh: the whole mesh on one GPU, or this rank's part of the library's own
    partition (onsas_part_*: every process builds the same partition and loads its rank).  Returns the context, the local ->
    global node map (None on one GPU), the number of owned nodes and of local tets."""
    if N == 1:
        ctx = ob.context_from_flat(mesh.xyz, tets=mesh.tets, mat_kind=kind, mat_params=params, free_dofs=free, device=local_rank)
        return ctx, None, mesh.n_nodes, mesh.n_tets
    from onsas_jl_b200 import multigpu
    P = ob.NativePartition(mesh.xyz, N, tets=mesh.tets, free_dofs=free)
    ctx = multigpu.make_distributed_context(P, kind, params, dist, local_rank, p2p=not no_p2p)
    l2g = P.local_to_global(rank).astype(np.int64)
    sz = P.sizes(rank)
    P.close()
    return ctx, l2g, sz["n_owned"], sz["n_tets"]


def state_error(ctx, mesh, l2g, n_own_nodes, U_ref_fn, dist):
    """max |U - U_ref| / max |U_ref| over the whole structure (owned nodes of every rank, all-reduced)."""
    import torch
    U = ctx.get_U().reshape(-1, 3)[:n_own_nodes]
    xyz = mesh.xyz if l2g is None else mesh.xyz[l2g[:n_own_nodes]]
    Ur = U_ref_fn(xyz).reshape(-1, 3)
    v = torch.tensor([np.abs(U - Ur).max(), np.abs(Ur).max()], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v[0] / v[1])


def residual_at(ctx, ob):
    """||(F_ext - F_int)[free]|| / ||F_ext|| of the state on the device, as the solver's prologue computes it (all ranks)."""
    info = ctx.step(ob.PRECOND_JACOBI, cg_maxiter=1, update_U=False)
    return info.norm_r / info.norm_Fext


def run_strong_c4(args, ob, N, rank, dist, local_rank, barrier, max_over_ranks):
    """configs[3]: synthetic structured SVK cube, 188^3 cells = 39 868 032 tetrahedra, the SAME mesh on 1 / 2 / 4 / 8 GPUs
    (strong scaling).  State: the last load step of examples/uniaxial_extension (E = 1, nu = 0.3, p = 3): the assembly is
    timed at the analytic equilibrium of load factor 1, the Newton step starts from the equilibrium of load factor 7/8."""
    import torch
    from onsas_jl_b200 import meshgen as mg
    n = args.c4_cells
    t0 = time.perf_counter()
    mesh = mg.box_tet_mesh(n, n, n, 1.0, 1.0, 1.0)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    p = 3.0
    a_prev, b_prev = uniaxial_state(_P_svk, p * 7.0 / 8.0, (1.8, 0.5))
    a_eq, b_eq = uniaxial_state(_P_svk, p, (1.9, 0.4))
    a_nw, b_nw = newton_iterate(_P_svk, p, a_prev, b_prev)
    Fext = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (p, 0.0, 0.0))
    t_mesh = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx, l2g, n_own, n_tets_local = make_context(ob, mesh, free, [ob.MAT_SVK], [[LAM, MU]], N, rank, dist, local_rank, args.no_p2p)
    t_setup = time.perf_counter() - t0
    stream = torch.cuda.Stream(device=local_rank)
    ctx.set_stream(stream.cuda_stream)
    loc = (lambda v: v) if l2g is None else (lambda v: v.reshape(-1, 3)[l2g].ravel())  # noqa: E731
    field = lambda a, b: (lambda xyz: mg.homogeneous_field(xyz, a, b))  # noqa: E731
    ctx.set_Fext(loc(Fext))
    ctx.set_U(loc(mg.homogeneous_field(mesh.xyz, a_eq, b_eq)))
    for _ in range(3):
        ctx.assemble()
    ctx.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    barrier()
    ev0.record(stream)
    for _ in range(K):
        ctx.assemble()
    ev1.record(stream)
    barrier()
    ms_asm = max_over_ranks(ev0.elapsed_time(ev1)) / K
    res_eq = residual_at(ctx, ob)        # parity (i): the analytic equilibrium state has no residual
    out = {"workload": f"configs[3]: synthetic structured SVK tet cube, {n}^3 cells = {mesh.n_tets} tets, {mesh.n_nodes * 3} dofs, "
                       "E = 1, nu = 0.3, p = 3 (examples/uniaxial_extension), partitioned over the GPUs by the library (RCB)",
           "scaling": "strong", "n_gpus": N, "n_tets": mesh.n_tets, "assembly_ms": ms_asm, "assembly_tets_per_s": mesh.n_tets / (ms_asm * 1e-3),
           "roofline_frac_assembly": ALGO_BYTES_PER_TET * n_tets_local / (ms_asm * 1e-3) / 1e9 / _peaks()[0],
           "residual_at_analytic_state": res_eq, "host_seconds": {"mesh": t_mesh, "partition_tables_upload": t_setup}}
    for name, pre in (("jacobi", ob.PRECOND_JACOBI), ("two_level", ob.PRECOND_TWO_LEVEL)):
        rec, ok = None, 1.0
        try:
            # untimed warm-up of this solver variant (two CG iterations): the first launch of a kernel pays its module load
            ctx.set_U(loc(mg.homogeneous_field(mesh.xyz, a_prev, b_prev)))
            ctx.newton_step(pre, cg_maxiter=2)
        except ob.OnsasError:
            pass
        ctx.set_U(loc(mg.homogeneous_field(mesh.xyz, a_prev, b_prev)))
        barrier()
        try:
            info = ctx.newton_step(pre)
            rec = {"ms": info.ms_assemble + info.ms_solve, "ms_assemble": info.ms_assemble, "ms_solve": info.ms_solve,
                   "cg_iters": int(info.cg_iters), "us_per_cg_iteration": 1e3 * info.ms_solve / max(int(info.cg_iters), 1)}
        except ob.OnsasError as exc:
            print(f"[bench] rank {rank}: C4 {name} Newton step failed: {exc}", file=sys.stderr)
            ok = 0.0
        ok = -max_over_ranks(-ok)
        ms_all = max_over_ranks(rec["ms"] if rec else 0.0)
        if ok > 0:
            rec["ms"] = ms_all
            # parity (ii): one exact Newton step from the homogeneous state (a_prev, b_prev) is the homogeneous state (a_nw, b_nw)
            rec["state_error_vs_analytic_newton_iterate"] = state_error(ctx, mesh, l2g, n_own, field(a_nw, b_nw), dist)
            out["newton_step_" + name] = rec
        else:
            out["newton_step_" + name] = None
    out["analytic"] = {"alpha_beta_start": [a_prev, b_prev], "alpha_beta_after_one_newton_step": [a_nw, b_nw], "alpha_beta_equilibrium": [a_eq, b_eq]}
    # ---- the complete solve of the target sentence at this size: 8 load steps from U = 0 with the reference's control flow
    #      (NonLinearStaticAnalyses.jl:70-104: load-step loop x Newton loop, isconverged! on the norms every rank receives --
    #      they are all-reduced inside the solver, so every rank takes the same decisions).  From 4 GPUs on by default
    #      (one GPU needs ~2 minutes for it).
    if N >= args.c4_full_solve_min_gpus:
        import math
        tols = ob.ConvergenceSettings(1e-8, 1e-8, 30)
        ctx.set_U(loc(np.zeros(mesh.n_nodes * 3)))
        iters, cgs = [], []
        barrier()
        t0 = time.perf_counter()
        for lam in np.linspace(1.0 / 8, 1.0, 8):
            ctx.set_Fext(loc(Fext * lam))
            it = ob.ResidualsIterationStep()
            cg = 0
            while isinstance(ob.isconverged(it, tols), ob.NotConvergedYet):
                info = ctx.newton_step(ob.PRECOND_TWO_LEVEL, 1e-10)
                it.update(info.norm_dU, info.norm_dU / info.norm_U if info.norm_U > 0 else math.inf,
                          info.norm_r, info.norm_r / info.norm_Fext if info.norm_Fext > 0 else math.inf)
                cg += int(info.cg_iters)
            iters.append(it.iter)
            cgs.append(cg)
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        out["full_solve"] = {"workload": "examples/uniaxial_extension on this mesh: 8 load steps from U = 0, tolerances 1e-8, two-level PCG at reltol 1e-10",
                             "wall_s": wall, "newton_iterations": iters, "reference_newton_iterations": [6, 5, 5, 4, 4, 4, 5, 5],
                             "cg_iterations_per_load_step": cgs,
                             "rel_error_vs_analytic_alpha2_beta_sqrt0.1": state_error(ctx, mesh, l2g, n_own, field(a_eq, b_eq), dist)}
    ctx.close()
    return out


def full_solve_leg(ob, device):
    """BASELINE.json's target sentence as one timed call: examples/uniaxial_extension (SVK, E = 1, nu = 0.3, p = 3, 2 x 1 x 1,
    NSTEPS = 8, tolerances 1e-8 -- uniaxial_extension.jl:11-24,116-120) on 88 x 44 x 44 cells = 1 022 208 tetrahedra from U = 0
    through solve(NonLinearStaticAnalysis(s; NSTEPS = 8), NewtonRaphson(tols)) of the mirrored API."""
    from onsas_jl_b200 import meshgen as mg
    mesh = mg.box_tet_mesh(88, 44, 44, 2.0, 1.0, 1.0)
    free = mg.free_dofs_from_fixed(mesh.n_nodes, 3, mg.uniaxial_fixed(mesh))
    unit = mg.global_face_load(mesh.n_nodes, mesh.xyz, mesh.faces["x1"], (3.0, 0.0, 0.0))
    s = ob.Structure.from_arrays(mesh.xyz, tets=mesh.tets, materials=[ob.SVK(E=1.0, nu=0.3)], free_dofs=free, fext=lambda t: unit * t)
    sa = ob.NonLinearStaticAnalysis(s, NSTEPS=8)
    nr = ob.NewtonRaphson(ob.ConvergenceSettings(1e-8, 1e-8, 30), preconditioner="two_level", cg_reltol=1e-10, device=device)
    sa.device_context(device)            # mesh upload + tables: outside the timed solve, as the reference builds its Structure first
    t0 = time.perf_counter()
    sol = ob.solve_(sa, nr)
    wall = time.perf_counter() - t0
    Ua = mg.homogeneous_field(mesh.xyz, 2.0, float(np.sqrt(0.1)))
    err = float(np.abs(sol.U[-1] - Ua).max() / np.abs(Ua).max())
    Rx = float(sol.reactions()[-1].reshape(-1, 3)[mesh.node_sets["x0"], 0].sum())
    sa._ctx.close()
    return {"workload": f"examples/uniaxial_extension SVK, 88 x 44 x 44 cells = {mesh.n_tets} tets, 8 load steps from U = 0, tol 1e-8",
            "wall_s": wall, "newton_iterations": sol.iterations(), "reference_newton_iterations": [6, 5, 5, 4, 4, 4, 5, 5],
            "cg_iterations_per_load_step": [int(sum(c)) for c in sol.cg_iterations],
            "rel_error_vs_analytic_alpha2_beta_sqrt0.1": err, "reaction_sum_x0": Rx, "expected_reaction": -3.0,
            "precond": "two_level", "cg_reltol": 1e-10}


def run_ours(args):
    import torch
    import onsas_jl_b200 as ob
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL logs to stdout by default: rank 0 prints ONE JSON line there
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = world
    assert N == args.gpus or world == 1, "--gpus must match the torchrun world size"

    mesh, free, U_half, U_prev, Fext = build_problem(args.cells, N)
    kind, params = [ob.MAT_NEOHOOKEAN], [[KBULK, MU]]
    n_tets_total = mesh.n_tets
    ctx, l2g, n_own_nodes, n_tets_local = make_context(ob, mesh, free, kind, params, N, rank, dist, local_rank, args.no_p2p)
    loc = (lambda v: v) if l2g is None else (lambda v: v.reshape(-1, 3)[l2g].ravel())  # noqa: E731  (global -> local, owned + halo)
    stream = torch.cuda.Stream(device=local_rank)   # the library launches on this stream; events are recorded on it
    ctx.set_stream(stream.cuda_stream)
    stats = ctx.table_stats()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, max(args.warmup, 3)
    ctx.set_U(loc(U_half))
    ctx.set_Fext(loc(Fext))
    extra_warmup = 0 if N == 1 else 10   # N > 1: the first grouped send/recv rounds between all neighbours are slow (channel set-up)
    for _ in range(W + extra_warmup):
        ctx.assemble()
    ctx.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- (1) device-resident assembly throughput: K launches of the fused assembly kernel
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        ctx.assemble()
    ev1.record(stream)
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ctx.synchronize()
    ms_step = ms_total / K
    value = n_tets_total / (ms_step * 1e-3)

    # ---- (2) end to end through the C ABI with pinned host buffers: U on the host in, F_int on the host out.
    #      onsas_assemble_host = assemble! with host state (H2D of U, kernel and D2H of F_int pipelined over slice ranges);
    #      the three separate calls (onsas_set_U + onsas_assemble + onsas_get_Fint) are timed beside it.
    hU, hF = pinned(ctx.n_dofs), pinned(ctx.n_dofs)
    hU[:] = loc(U_half)
    lib, h = ctx._lib, ctx._h
    for _ in range(3):
        assert lib.onsas_assemble_host(h, hU, hF) == 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        assert lib.onsas_assemble_host(h, hU, hF) == 0
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    e2e_value = n_tets_total / (ms_e2e * 1e-3)
    F_pipe = np.array(hF, copy=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        assert lib.onsas_set_U(h, hU) == 0
        assert lib.onsas_assemble(h) == 0
        assert lib.onsas_get_Fint(h, hF) == 0
    barrier()
    ms_e2e3 = max_over_ranks((time.perf_counter() - t0) * 1e3) / K
    assert np.array_equal(F_pipe, np.asarray(hF)), "onsas_assemble_host and the three-call path disagree"

    # ---- (3) Newton-step time: assemble! + step! (Jacobi-PCG at the reference's default tolerance sqrt(eps))
    newton = []
    for _ in range(max(1, min(K, 3))):
        ctx.set_U(loc(U_prev))
        barrier()
        info = ctx.newton_step(ob.PRECOND_JACOBI)
        newton.append((max_over_ranks(info.ms_assemble + info.ms_solve), info.ms_assemble, info.ms_solve, int(info.cg_iters),
                       info.norm_r / info.norm_Fext, info.norm_dU / max(info.norm_U, 1e-300)))
    newton.sort()
    nw = newton[len(newton) // 2]
    # parity carried by the line itself (every N): (i) the analytic equilibrium state of the load has no residual; (ii) one
    # exact Newton step from the homogeneous state U_prev lands on the homogeneous state of the scalar Newton iterate
    from onsas_jl_b200 import meshgen as mg
    a_nw, b_nw = newton_iterate(_P_neo, -0.5, *_neo_state(4.0 / 9.0))
    err_newton = state_error(ctx, mesh, l2g, n_own_nodes, lambda xyz: mg.homogeneous_field(xyz, a_nw, b_nw), dist)
    ctx.set_U(loc(U_half))
    ctx.assemble()
    res_eq = residual_at(ctx, ob)
    # ---- (3b) the same Newton step with the two-level preconditioner (Jacobi + aggregated coarse space);
    #      its time includes the coarse set-up (E = Z^T K Z and its inverse are rebuilt after every assembly)
    #      Every rank always takes part in the collectives below, whatever happened in its own solve.
    nw2, two_level_ok = None, 1.0
    newton2 = []
    for _ in range(max(1, min(K, 3))):
        ctx.set_U(loc(U_prev))
        barrier()
        ms_local, rec = float("nan"), None
        if two_level_ok > 0:
            try:
                info = ctx.newton_step(ob.PRECOND_TWO_LEVEL)
                ms_local = info.ms_assemble + info.ms_solve
                rec = (info.ms_assemble, info.ms_solve, int(info.cg_iters), info.norm_dU / max(info.norm_U, 1e-300))
            except ob.OnsasError as exc:   # e.g. --no-p2p: only the streamed persistent solver implements it
                print(f"[bench] rank {rank}: two-level Newton step skipped: {exc}", file=sys.stderr)
        two_level_ok = -max_over_ranks(-(1.0 if rec is not None else 0.0))   # min over ranks: did every rank succeed?
        ms_all = max_over_ranks(ms_local if rec is not None else 0.0)
        if two_level_ok > 0:
            newton2.append((ms_all,) + rec)
    if two_level_ok > 0 and newton2:
        newton2.sort()
        nw2 = newton2[len(newton2) // 2]

    # ---- (4) SpMV alone (secondary roofline)
    for _ in range(3):
        ctx.spmv_resident()
    barrier()
    ev0.record(stream)
    for _ in range(20):
        ctx.spmv_resident()
    ev1.record(stream)
    barrier()
    ms_spmv = max_over_ranks(ev0.elapsed_time(ev1)) / 20
    clocks = sampler.stop() if sampler else None
    stats_w, n_dofs_w, n_owned_w = stats, ctx.n_dofs, ctx.n_owned
    ctx.close()
    del ctx

    # ---- (5) the full target solve at size (every N): examples/uniaxial_extension, 8 load steps from U = 0, through the
    #      mirrored reference API on this rank's device(s) -- N = 1 only here (the API drives its devices from one process)
    # (the two extra legs must never cost the headline line: a failure is reported inside the line instead)
    full = None
    if N == 1 and not args.no_full_solve:
        try:
            full = full_solve_leg(ob, local_rank)
        except Exception as exc:  # noqa: BLE001
            full = {"error": repr(exc)[:300]}
    # ---- (6) configs[3] on the same N GPUs, strong scaling
    c4 = None
    if not args.no_c4:
        try:
            c4 = run_strong_c4(args, ob, N, rank, dist, local_rank, barrier, max_over_ranks)
        except Exception as exc:  # noqa: BLE001
            c4 = {"error": repr(exc)[:300]}
            print(f"[bench] rank {rank}: configs[3] leg failed: {exc!r}", file=sys.stderr)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    achieved = ALGO_BYTES_PER_TET * n_tets_local / (ms_step * 1e-3) / 1e9   # per GPU: its own tets per its launch
    traffic = None   # DRAM bytes per launch of the assembly kernel from the committed ncu capture of the same workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_assemble"]
        if tr["n_tets"] == n_tets_local:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    nnzb = stats["nnz_blocks"]
    n_own_dofs = n_owned_w * 3
    spmv_bytes = 72 * nnzb + 4 * nnzb + 8 * (stats["n_slices"] + 1) + 16 * n_own_dofs + n_own_dofs  # BSR-3x3 + mask
    out = {
        "metric": "tet_fint_Kt_assembled_elements_per_s", "value": value, "unit": "tets/s", "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.cells, N, n_tets_total, mesh.n_nodes * 3),
        "implementation": {"extra_warmup_steps": extra_warmup,
                           "partition": "single GPU" if N == 1 else f"RCB slabs, {N} ranks (library partitioner), no exchange inside the assembly (U carries its halo part); CG: " +
                                        ("one launch per phase, NCCL between them" if args.no_p2p else
                                         "persistent TMA-streamed kernel, halo + all-reduce pushed over NVLink peer memory")},
        "newton_step_ms": nw[0], "newton_step": {"ms_assemble": nw[1], "ms_solve": nw[2], "cg_iters": nw[3], "precond": "jacobi",
                                                  "cg_reltol": float(np.sqrt(np.finfo(np.float64).eps)), "rel_residual_in": nw[4],
                                                  "rel_dU": nw[5]},
        "newton_step_two_level": None if nw2 is None else {
            "ms": nw2[0], "ms_assemble": nw2[1], "ms_solve": nw2[2], "cg_iters": nw2[3], "rel_dU": nw2[4],
            "precond": "jacobi + aggregated coarse space (precond = 2), coarse set-up inside ms_solve"},
        "e2e": {"value": e2e_value, "unit": "tets/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(n_dofs_w * 8),
                "d2h_bytes_per_step": int(n_dofs_w * 8),
                "what": "onsas_assemble_host(pinned U in, pinned F_int out): H2D, kernel and D2H pipelined over slice ranges",
                "three_calls_ms_per_step": ms_e2e3, "three_calls_value": n_tets_total / (ms_e2e3 * 1e-3)},
        "gpu_launches": K,   # timed device-resident region: one fused assembly kernel per step (no halo exchange: U arrives with its halo part); the e2e region launches one kernel per slice range
        "roofline": {"bound": "hbm", "kernel": "k_assemble<tet,NeoHookean>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_tet": ALGO_BYTES_PER_TET,
                     "note": "algorithmic bytes count K_e/f_e once per element; the fused kernel writes each K entry once "
                             "(compulsory DRAM traffic ~0.45 KB/tet), so frac > 1 is on-chip reuse, not extra bandwidth"},
        "roofline_spmv": {"bound": "hbm", "kernel": "k_spmv_dot<3>", "achieved": spmv_bytes / (ms_spmv * 1e-3) / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": spmv_bytes / (ms_spmv * 1e-3) / 1e9 / peak, "ms": ms_spmv,
                          "bytes": int(spmv_bytes)},
        # one PCG iteration of the persistent solver: SpMV bytes above + 10 vector passes, over the
        # measured time per iteration of the Newton step's solve
        "roofline_pcg": {"bound": "hbm", "kernel": "cg_stream<3> (one CG iteration)", "bytes": int(spmv_bytes + 80 * n_own_dofs),
                         "us_per_iteration": 1e3 * nw[2] / max(nw[3], 1),
                         "achieved": (spmv_bytes + 80 * n_own_dofs) / (1e-3 * nw[2] / max(nw[3], 1)) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": (spmv_bytes + 80 * n_own_dofs) / (1e-3 * nw[2] / max(nw[3], 1)) / 1e9 / peak},
        "parity": {"residual_at_analytic_state": res_eq, "state_error_vs_analytic_newton_iterate": err_newton,
                   "what": "relative residual of the analytic equilibrium state; max |U - U_analytic| / max |U| after the timed Jacobi Newton step "
                           "(one exact Newton step from a homogeneous state is the homogeneous state of the scalar Newton iterate)"},
        "full_solve": full, "strong_c4": c4,
        "tables": stats, "clocks": clocks,
    }
    if N == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.cells, threads=1)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(cells: int, threads: int, reps: int = 1):
    """The oracle (CPU restatement of the reference's algorithm: serial element loop, COO triplets, per-entry sparse
    insertion) timed on the host cores on the same mesh.  kind = "port": the reference is Julia and cannot run here."""
    from oracle import oracle as O
    mesh, free, U_half, _, _ = build_problem(cells, 1)
    m = O.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[O.MAT_NEOHOOKEAN], mat_params=[[KBULK, MU]], free_dofs=free)
    if threads == 1:
        os.environ["OMP_NUM_THREADS"] = "1"
        asm = O.Assembly(m)
    else:
        asm = O.AssemblyMT(m)
    asm.assemble(U_half)  # warm-up (page faults, pattern)
    t0 = time.perf_counter()
    for _ in range(reps):
        asm.assemble(U_half)
    dt = (time.perf_counter() - t0) / reps
    return {"value": mesh.n_tets / dt, "unit": "tets/s", "cores": threads if threads == 1 else O.lib().orc_num_threads(),
            "kind": "port", "seconds_per_pass": dt,
            "sample": f"{reps} full assembly pass(es) of the same {mesh.n_tets}-tet mesh (oracle/onsas_oracle.c, gcc -O2)"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on the host cores (oracle port, all threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    cells = min(args.cells, 55)
    from oracle import oracle as O
    mesh, free, U_half, _, _ = build_problem(cells, 1)
    m = O.FlatModel(xyz=mesh.xyz, tets=mesh.tets, mat_kind=[O.MAT_NEOHOOKEAN], mat_params=[[KBULK, MU]], free_dofs=free)
    try:   # torchrun sets OMP_NUM_THREADS=1 for its workers: the reference arm uses all host cores
        O.lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    asm = O.AssemblyMT(m)
    for _ in range(max(1, min(W, 2))):
        asm.assemble(U_half)
    Kb = 0
    t0 = time.perf_counter()
    while Kb < K and (Kb == 0 or time.perf_counter() - t0 < 120.0):   # all K steps unless the host is so slow that they would take minutes
        asm.assemble(U_half)
        Kb += 1
    dt = (time.perf_counter() - t0) / Kb
    val = mesh.n_tets / dt
    cores = O.lib().orc_num_threads()
    out = {"impl": "reference", "metric": "tet_fint_Kt_assembled_elements_per_s", "value": val, "unit": "tets/s",
           "n_gpus": args.gpus, "steps": Kb, "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(cells, 1, mesh.n_tets, mesh.n_nodes * 3),   # the same keys and strings as our arm's line at N = 1
           "implementation": {"what": "the reference's algorithm restated in C (ONSAS.jl is Julia; no Julia toolchain in this image), OpenMP on all host cores",
                              "note": "the CPU arm always runs the 1-GPU mesh (rates are compared: tets/s); at --gpus N > 1 our arm's mesh is N times larger"},
           "cpu_baseline": {"value": val, "unit": "tets/s", "cores": cores, "kind": "port",
                            "sample": f"{Kb} full assembly passes, OpenMP over elements + row-parallel gather"},
           "e2e": {"value": val, "unit": "tets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=55, help="hexes per edge per GPU (55 -> 998 250 tets)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL multi-launch CG instead of the peer-memory persistent kernel")
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[3] strong-scaling leg (39.9 M tets)")
    ap.add_argument("--c4-cells", type=int, default=188, help="cells per edge of the configs[3] cube (188 -> 39 868 032 tets)")
    ap.add_argument("--no-full-solve", action="store_true", help="skip the full 8-load-step solve leg")
    ap.add_argument("--c4-full-solve-min-gpus", type=int, default=4, help="run the complete 8-load-step solve on the configs[3] mesh from this many GPUs on")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
