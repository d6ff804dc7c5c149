"""Element-wise domain decomposition for the multi-GPU path: one process per GPU,
each rank OWNS a contiguous range of (renumbered) nodes = block rows of K, evaluates every element that
touches an owned node (interface elements are evaluated on both sides, no assembly communication), and
needs the values of its halo nodes (the other nodes of those elements) for the element gathers and the SpMV.

The reference has no distributed path at all (single process, StaticAnalyses.jl:105-118); this is new.
Host-side numpy only -- the exchange itself is done by libonsas_cuda (NCCL send/recv over NVLink).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def rcb_order(xyz: np.ndarray, n_parts: int):
    """Recursive coordinate bisection of the nodes.  Returns (order, ranges): `order[k]` = old id of the node that
    becomes new id k; part p owns new ids ranges[p] .. ranges[p+1].  Inside a part the original order is kept
    (structured meshes stay x-fastest, which keeps the row-owner pairs and SpMV gathers local)."""
    n = xyz.shape[0]
    parts = [np.arange(n, dtype=np.int64)]
    counts = [n_parts]
    while any(c > 1 for c in counts):
        new_parts, new_counts = [], []
        for ids, c in zip(parts, counts):
            if c == 1:
                new_parts.append(ids)
                new_counts.append(1)
                continue
            cl = c // 2
            pts = xyz[ids]
            axis = int(np.argmax(pts.max(axis=0) - pts.min(axis=0))) if len(ids) else 0
            k = (len(ids) * cl) // c
            # stable split at the k-th coordinate along the axis (ties broken by node id)
            key = np.lexsort((ids, pts[:, axis]))
            left = np.sort(ids[key[:k]])
            right = np.sort(ids[key[k:]])
            new_parts += [left, right]
            new_counts += [cl, c - cl]
        parts, counts = new_parts, new_counts
    order = np.concatenate(parts) if parts else np.zeros(0, np.int64)
    ranges = np.zeros(n_parts + 1, np.int64)
    ranges[1:] = np.cumsum([len(p) for p in parts])
    return order, ranges


@dataclass
class LocalPart:
    rank: int
    n_ranks: int
    n_owned: int
    local_to_global: np.ndarray   # (n_local,) global node id of each local node: owned first, then halo by owner
    xyz: np.ndarray               # (n_local, dim)
    tets: np.ndarray              # (n, 4) local ids
    tet_global: np.ndarray        # global element id of each local tet
    tet_mat: np.ndarray
    trusses: np.ndarray
    truss_global: np.ndarray
    truss_mat: np.ndarray
    truss_area: np.ndarray
    free_dofs: np.ndarray         # free local dofs: of the owned nodes first, then of the halo nodes
    n_free_global: int
    nbr_rank: np.ndarray          # neighbours, ascending
    send_ptr: np.ndarray
    send_nodes: np.ndarray        # local (owned) node ids, grouped by neighbour, ascending global id
    recv_ptr: np.ndarray          # halo offsets (relative to n_owned), grouped by neighbour

    @property
    def n_local(self):
        return len(self.local_to_global)

    def scatter_global(self, v_global: np.ndarray, dim: int) -> np.ndarray:
        """global dof vector -> local (owned + halo) dof vector"""
        return v_global.reshape(-1, dim)[self.local_to_global].ravel()

    def owned_global_dofs(self, dim: int) -> np.ndarray:
        g = self.local_to_global[:self.n_owned]
        return (g[:, None] * dim + np.arange(dim)[None, :]).ravel()


def build_local_part(rank: int, ranges: np.ndarray, xyz: np.ndarray, tets=None, tet_mat=None, trusses=None,
                     truss_mat=None, truss_area=None, free_dofs=None) -> LocalPart:
    """Everything rank `rank` uploads, from the GLOBAL mesh in the partition numbering
    (node ids already permuted so that rank p owns ids ranges[p]..ranges[p+1])."""
    n_ranks = len(ranges) - 1
    dim = xyz.shape[1]
    lo, hi = int(ranges[rank]), int(ranges[rank + 1])
    tets = np.zeros((0, 4), np.int32) if tets is None else np.asarray(tets)
    trusses = np.zeros((0, 2), np.int32) if trusses is None else np.asarray(trusses)

    def touching(conn):
        if len(conn) == 0:
            return np.zeros(0, np.int64)
        owned = (conn >= lo) & (conn < hi)
        return np.nonzero(owned.any(axis=1))[0]

    te, be = touching(tets), touching(trusses)
    used = np.unique(np.concatenate([tets[te].ravel(), trusses[be].ravel(), np.arange(lo, hi)]))
    halo = used[(used < lo) | (used >= hi)]
    owner = np.searchsorted(ranges, halo, side="right") - 1
    key = np.lexsort((halo, owner))
    halo, owner = halo[key], owner[key]
    l2g = np.concatenate([np.arange(lo, hi, dtype=np.int64), halo.astype(np.int64)])
    g2l = {}
    lut = np.full(xyz.shape[0], -1, np.int64)
    lut[l2g] = np.arange(len(l2g))
    nbr = np.unique(owner)
    recv_ptr = np.zeros(len(nbr) + 1, np.int64)
    for k, r in enumerate(nbr):
        recv_ptr[k + 1] = recv_ptr[k] + np.count_nonzero(owner == r)
    # what each neighbour needs from me: my owned nodes that share an element with one of ITS owned nodes
    send_lists = []
    for r in nbr:
        rlo, rhi = int(ranges[r]), int(ranges[r + 1])
        need = []
        for conn in (tets[te], trusses[be]):
            if len(conn) == 0:
                continue
            theirs = ((conn >= rlo) & (conn < rhi)).any(axis=1)
            c = conn[theirs].ravel()
            need.append(c[(c >= lo) & (c < hi)])
        g = np.unique(np.concatenate(need)) if need else np.zeros(0, np.int64)
        send_lists.append(lut[g])
    send_ptr = np.zeros(len(nbr) + 1, np.int64)
    for k, s in enumerate(send_lists):
        send_ptr[k + 1] = send_ptr[k] + len(s)
    send_nodes = np.concatenate(send_lists).astype(np.int32) if send_lists else np.zeros(0, np.int32)

    if free_dofs is None:
        free_dofs = np.arange(xyz.shape[0] * dim, dtype=np.int64)
    free_dofs = np.asarray(free_dofs, np.int64)
    fmask = np.zeros(xyz.shape[0] * dim, bool)
    fmask[free_dofs] = True
    own_free = np.nonzero(fmask[lo * dim: hi * dim])[0].astype(np.int64)  # local dof = global dof - lo*dim
    halo_free = np.nonzero(fmask.reshape(-1, dim)[halo].ravel())[0].astype(np.int64) + (hi - lo) * dim
    own_free = np.concatenate([own_free, halo_free])   # halo dofs after the owned ones: the solver updates U there too
    return LocalPart(
        rank=rank, n_ranks=n_ranks, n_owned=hi - lo, local_to_global=l2g, xyz=np.ascontiguousarray(xyz[l2g]),
        tets=lut[tets[te]].astype(np.int32).reshape(-1, 4), tet_global=te,
        tet_mat=(np.zeros(len(te), np.int32) if tet_mat is None else np.asarray(tet_mat, np.int32)[te]),
        trusses=lut[trusses[be]].astype(np.int32).reshape(-1, 2), truss_global=be,
        truss_mat=(np.zeros(len(be), np.int32) if truss_mat is None else np.asarray(truss_mat, np.int32)[be]),
        truss_area=(np.ones(len(be)) if truss_area is None else np.asarray(truss_area, np.float64)[be]),
        free_dofs=own_free, n_free_global=len(free_dofs), nbr_rank=nbr.astype(np.int32), send_ptr=send_ptr,
        send_nodes=send_nodes, recv_ptr=recv_ptr)


def renumber(order: np.ndarray, xyz: np.ndarray, *conns):
    """Apply a node permutation (order[new] = old) to coordinates and connectivity arrays."""
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    return (xyz[order],) + tuple(None if c is None else inv[np.asarray(c)].astype(np.int32) for c in conns) + (inv,)
