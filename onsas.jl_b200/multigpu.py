"""One-process-per-GPU setup of a distributed DeviceContext: NCCL communicator bootstrap
through torch.distributed (plumbing), halo plan upload, and the CUDA-IPC window exchange that lets the persistent
CG kernel push halo values / partial sums straight into the peers' memory over NVLink."""
from __future__ import annotations

import numpy as np

from .device import DeviceContext


def make_distributed_context(part, mat_kind, mat_params, dist, local_rank: int, p2p: bool = True,
                             truss_strain: int = 0) -> DeviceContext:
    """`part` is a partition.LocalPart (the numpy restatement) or a device.NativePartition (the library's partitioner:
    every process built the same partition and loads its own rank); `dist` an initialised torch.distributed (NCCL)."""
    import torch
    from .device import NativePartition
    rank, world = dist.get_rank(), dist.get_world_size()
    native = isinstance(part, NativePartition)
    ctx = DeviceContext(local_rank)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(DeviceContext.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    if native:
        ctx.set_materials(mat_kind, mat_params)
        ctx.load_part(part, rank, truss_strain)
    else:
        ctx.set_nodes(part.xyz, part.n_owned)
        ctx.set_materials(mat_kind, mat_params)
        if len(part.tets):
            ctx.set_tets(part.tets, part.tet_mat)
        if len(part.trusses):
            ctx.set_trusses(part.trusses, part.truss_area, part.truss_mat, truss_strain)
        ctx.set_free_dofs(part.free_dofs, part.n_free_global)
        ctx.set_halo(part.nbr_rank, part.send_ptr, part.send_nodes, part.recv_ptr)
    ctx.finalize()
    if p2p and world > 1:
        # every rank must take the same path: export first, agree on success, then import
        try:
            handle, offset = ctx.p2p_export()
            ok = 1
        except Exception as ex:  # e.g. CUDA IPC not permitted: the NCCL per-phase CG path still runs on the GPUs
            handle, offset, ok = b"", 0, 0
            print(f"[onsas] rank {rank}: peer-memory export failed ({ex}); using the NCCL CG path", flush=True)
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return ctx
        meta = [None] * world
        if native:   # the partition knows where this rank's values start inside every neighbour's halo
            dist.all_gather_object(meta, dict(handle=handle, offset=offset))
            ctx.p2p_import([m["handle"] for m in meta], [m["offset"] for m in meta], None)
            return ctx
        dist.all_gather_object(meta, dict(handle=handle, offset=offset, n_owned=int(part.n_owned),
                                          nbr=[int(r) for r in part.nbr_rank], recv_ptr=[int(v) for v in part.recv_ptr]))
        remote = []
        for r in part.nbr_rank:          # where my values start inside neighbour r's halo receive buffer
            m = meta[int(r)]
            j = m["nbr"].index(rank)
            remote.append(m["recv_ptr"][j])
        ctx.p2p_import([m["handle"] for m in meta], [m["offset"] for m in meta], np.asarray(remote, np.int64))
    return ctx
