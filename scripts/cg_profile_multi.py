"""Per-phase SM-clock cycles of the peer-memory persistent CG on N GPUs (one process per GPU) and the NVLink traffic of one Newton
step, to attribute the per-iteration gap between 1 and N GPUs (poll latency vs barrier skew vs the halo push):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 scripts/cg_profile_multi.py
Weak-scaling problem of bench.py (55^3 cells per GPU, NeoHookean).  Rank 0 prints one JSON line per solver variant with every
rank's cycles per CG iteration (block 0 of each rank) and the NVLink byte counters (nvidia-smi nvlink -gt d) around the step."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import onsas_jl_b200 as ob  # noqa: E402

L = ob._lib
world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


def nvlink_bytes():
    """sum over links of (tx, rx) KiB of GPU 0 (nvidia-smi nvlink -gt d), or None"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], capture_output=True, text=True, timeout=20).stdout
        tx = sum(int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
        rx = sum(int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
        nvlink_bytes.raw = out[:600]
        return tx, rx
    except Exception:
        return None


mesh, free, U_half, U_prev, Fext = bench.build_problem(55, world)
ctx, l2g, n_own, n_tets_local = bench.make_context(ob, mesh, free, [ob.MAT_NEOHOOKEAN], [[bench.KBULK, bench.MU]], world, rank, dist, local_rank)
loc = (lambda v: v) if l2g is None else (lambda v: v.reshape(-1, 3)[l2g].ravel())
ctx.set_Fext(loc(Fext))
for name, pre, sr, glob in (("jacobi single-reduction", ob.PRECOND_JACOBI, 1, 0), ("jacobi classic", ob.PRECOND_JACOBI, 0, 0), ("two-level", ob.PRECOND_TWO_LEVEL, 0, 0),
                            ("two-level + global coarse level", ob.PRECOND_TWO_LEVEL, 0, 1)):
    ctx.set_option(L.OPT_CG_SINGLE_REDUCTION, sr)
    ctx.set_option(L.OPT_COARSE_GLOBAL, glob)
    if glob:
        ctx.set_U(loc(U_prev))
        ctx.newton_step(pre, cg_maxiter=2)             # the first launch of the set-up kernels pays their module load
    ctx.set_option(L.OPT_CG_PROFILE, 0)
    ctx.set_U(loc(U_prev))
    info0 = ctx.newton_step(pre)                       # un-profiled timing
    if dist is not None:
        dist.barrier()
    nv0 = nvlink_bytes() if rank == 0 else None
    ctx.set_option(L.OPT_CG_PROFILE, 1)
    ctx.set_U(loc(U_prev))
    info = ctx.newton_step(pre)
    if dist is not None:
        dist.barrier()
    nv1 = nvlink_bytes() if rank == 0 else None
    pv = ctx.cg_profile()
    pv.pop("slowest_cta_spmv", 0)
    keys = list(pv.keys())
    t = torch.tensor([pv[k] / max(int(info.cg_iters), 1) for k in keys] + [info0.ms_solve, info.ms_solve], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(t) for _ in range(world)]
    if dist is not None:
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        rows = [[round(float(x)) for x in v[:len(keys)]] for v in allv]
        halo_dofs = 0 if l2g is None else 3 * (len(l2g) - n_own)
        rec = {"variant": name, "n_gpus": world, "cg_iters": int(info.cg_iters), "us_per_iteration": 1e3 * float(max(v[-2] for v in allv)) / max(int(info0.cg_iters), 1),
               "us_per_iteration_profiled": 1e3 * float(max(v[-1] for v in allv)) / max(int(info.cg_iters), 1),
               "phases": keys, "cycles_per_iteration_by_rank": rows, "halo_dofs_rank0": halo_dofs,
               "ll_bytes_pushed_per_iteration_rank0_estimate": 16 * halo_dofs + 16 * 4 * world}
        if nv0 and nv1:
            rec["nvlink_gpu0_KiB_during_the_profiled_step"] = {"tx": nv1[0] - nv0[0], "rx": nv1[1] - nv0[1]}
            rec["nvlink_gpu0_bytes_per_iteration"] = {"tx": 1024.0 * (nv1[0] - nv0[0]) / max(int(info.cg_iters), 1), "rx": 1024.0 * (nv1[1] - nv0[1]) / max(int(info.cg_iters), 1)}
        rec["nvidia_smi_nvlink_raw_head"] = getattr(nvlink_bytes, "raw", "")[:300]
        print(json.dumps(rec), flush=True)
ctx.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
