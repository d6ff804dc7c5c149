"""Thin object wrapper over the C ABI: one `DeviceContext` = one `onsas_ctx` = the device-resident
analogue of the reference's `FullStaticState` (StructuralAnalyses/StaticStates.jl:33-102).

Everything that computes goes through libonsas_cuda (CUDA, sm_100a).  No CPU path exists here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class OnsasError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libonsas_cuda status {status}: {message}")
        self.status = status


class NegativeVolumeError(ValueError):
    """ArgumentError("Element with negative volume, check connectivity.") of Tetrahedrons.jl:136."""


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceContext:
    def __init__(self, device=0):
        """`device`: one CUDA device index, or a sequence of them -- then ONE host process drives all of them behind the
        same interface (onsas_create_multi): global mesh and global vectors in, partitioning / halo plan / peer-memory
        wiring inside onsas_finalize_mesh."""
        self._lib = L.lib()
        h = C.c_void_p()
        if np.ndim(device) > 0:
            devs = _as(device, np.int32).ravel()
            st = self._lib.onsas_create_multi(devs, len(devs), C.byref(h))
            self.devices = [int(d) for d in devs]
        else:
            st = self._lib.onsas_create(int(device), C.byref(h))
            self.devices = [int(device)]
        if st != L.OK:
            raise OnsasError(st, (self._lib.onsas_last_error(None) or b"").decode())
        self._h = h
        self.dim = 3
        self.n_nodes = 0
        self.n_owned = 0
        self.n_tets = 0
        self.n_trusses = 0

    # -- plumbing
    def _check(self, st: int):
        if st == L.OK:
            return
        msg = (self._lib.onsas_last_error(self._h) or b"").decode()
        if st == L.ERR_NEGATIVE_VOLUME:
            raise NegativeVolumeError(msg)
        raise OnsasError(st, msg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.onsas_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int | None):
        self._check(self._lib.onsas_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_option(self, key: int, value: int):
        self._check(self._lib.onsas_set_option(self._h, key, int(value)))

    # -- mesh
    def set_nodes(self, xyz, n_owned: int | None = None):
        xyz = _as(xyz, np.float64)
        if xyz.ndim == 1:
            xyz = xyz.reshape(-1, 1)
        self.n_nodes, self.dim = xyz.shape
        self.n_owned = self.n_nodes if n_owned is None else int(n_owned)
        self._check(self._lib.onsas_set_nodes(self._h, self.n_nodes, self.n_owned, self.dim, xyz.ravel()))

    def set_materials(self, kind, params):
        kind = _as(kind, np.int32).ravel()
        params = _as(params, np.float64).reshape(-1, 2)
        self._check(self._lib.onsas_set_materials(self._h, len(kind), kind, params.ravel()))

    def set_tets(self, conn, mat_id=None):
        conn = _as(conn, np.int32).reshape(-1, 4)
        mat = None if mat_id is None else _as(mat_id, np.int32)
        self.n_tets = len(conn)
        self._check(self._lib.onsas_set_tets(self._h, len(conn), _ptr(conn), _ptr(mat)))

    def set_trusses(self, conn, area, mat_id=None, strain_model: int = L.STRAIN_ROTATED_ENGINEERING):
        conn = _as(conn, np.int32).reshape(-1, 2)
        area = _as(area, np.float64).ravel()
        assert len(area) == len(conn)
        mat = None if mat_id is None else _as(mat_id, np.int32)
        self.n_trusses = len(conn)
        self._check(self._lib.onsas_set_trusses(self._h, len(conn), _ptr(conn), _ptr(mat), _ptr(area), int(strain_model)))

    def set_free_dofs(self, free_dofs, n_free_global: int = 0):
        fd = _as(free_dofs, np.int64).ravel()
        self._check(self._lib.onsas_set_free_dofs(self._h, len(fd), fd, int(n_free_global)))

    def finalize(self):
        self._check(self._lib.onsas_finalize_mesh(self._h))

    @property
    def n_dofs(self):
        return self.n_nodes * self.dim

    # -- state
    def set_U(self, U):
        U = _as(U, np.float64).ravel()
        assert U.size == self.n_dofs
        self._check(self._lib.onsas_set_U(self._h, U))

    def get_U(self):
        out = np.empty(self.n_dofs)
        self._check(self._lib.onsas_get_U(self._h, out))
        return out

    def set_Fext(self, F):
        F = _as(F, np.float64).ravel()
        assert F.size == self.n_dofs
        self._check(self._lib.onsas_set_Fext(self._h, F))

    # -- external loads built on the device (apply!, StructuralAnalyses.jl:228-241)
    def add_face_load(self, tri, kind: int, values) -> int:
        """kind 0: GlobalLoad values[3] * A/3 per face node; kind 1: Pressure -n * A/3 * values[0]."""
        tri = _as(tri, np.int32).reshape(-1, 3)
        v = np.zeros(3)
        v[:np.size(values)] = np.ravel(values)
        pid = C.c_int32(-1)
        self._check(self._lib.onsas_add_face_load(self._h, len(tri), tri.ravel() if len(tri) else np.zeros(3, np.int32), int(kind), v, C.byref(pid)))
        return int(pid.value)

    def add_nodal_load(self, nodes, values) -> int:
        nodes = _as(nodes, np.int32).ravel()
        v = np.zeros(3)
        v[:np.size(values)] = np.ravel(values)
        pid = C.c_int32(-1)
        self._check(self._lib.onsas_add_nodal_load(self._h, len(nodes), nodes if len(nodes) else np.zeros(1, np.int32), v, C.byref(pid)))
        return int(pid.value)

    def apply_loads(self, factors):
        f = _as(factors, np.float64).ravel()
        self._check(self._lib.onsas_apply_loads(self._h, len(f), f if len(f) else np.zeros(1)))

    def clear_loads(self):
        self._check(self._lib.onsas_clear_loads(self._h))

    def get_Fext(self):
        out = np.empty(self.n_dofs)
        self._check(self._lib.onsas_get_Fext(self._h, out))
        return out

    def get_Fint(self):
        out = np.empty(self.n_dofs)
        self._check(self._lib.onsas_get_Fint(self._h, out))
        return out

    def get_dU(self):
        out = np.empty(self.n_dofs)
        self._check(self._lib.onsas_get_dU(self._h, out))
        return out

    # -- hot path
    def assemble(self):
        """assemble!(s, sa) -- asynchronous; errors surface at the next synchronizing call."""
        self._check(self._lib.onsas_assemble(self._h))

    def assemble_host(self, U, out=None):
        """assemble!(s, sa) with host state on both sides: U in, F_int out, the copies pipelined with the kernel
        (onsas_assemble_host).  `out` may be a preallocated (pinned) float64 array of n_dofs."""
        U = _as(U, np.float64).ravel()
        assert U.size == self.n_dofs
        out = np.empty(self.n_dofs) if out is None else out
        self._check(self._lib.onsas_assemble_host(self._h, U, out))
        return out

    def synchronize(self):
        self._check(self._lib.onsas_synchronize(self._h))

    def eval_elements(self, family: int = L.FAMILY_TET, first: int = 0, count: int | None = None):
        ne = self.n_tets if family == L.FAMILY_TET else self.n_trusses
        count = ne - first if count is None else count
        nde = (4 if family == L.FAMILY_TET else 2) * self.dim
        f = np.empty((count, nde))
        K = np.empty((count, nde * nde))
        s = np.empty((count, 9))
        e = np.empty((count, 9))
        self._check(self._lib.onsas_eval_elements(self._h, family, first, count, f.ravel(), K.ravel(), s.ravel(), e.ravel()))
        return f, K, s, e

    def newton_step(self, precond=L.PRECOND_JACOBI, cg_reltol=None, cg_abstol=0.0, cg_maxiter=0) -> L.StepInfo:
        if cg_reltol is None:
            cg_reltol = float(np.sqrt(np.finfo(np.float64).eps))  # StructuralSolvers.jl:229-234
        info = L.StepInfo()
        self._check(self._lib.onsas_newton_step(self._h, precond, cg_reltol, cg_abstol, cg_maxiter, C.byref(info)))
        return info

    def step(self, precond=L.PRECOND_JACOBI, cg_reltol=None, cg_abstol=0.0, cg_maxiter=0, update_U=True) -> L.StepInfo:
        if cg_reltol is None:
            cg_reltol = float(np.sqrt(np.finfo(np.float64).eps))
        info = L.StepInfo()
        # update_U: False / 0 = leave U, True / 1 = U[free] += dU, 2 = the linear-analysis step (r = F_ext[free], U[free] = dU)
        self._check(self._lib.onsas_step(self._h, precond, cg_reltol, cg_abstol, cg_maxiter, int(update_U), C.byref(info)))
        return info

    def pcg(self, b, precond=L.PRECOND_JACOBI, reltol=None, abstol=0.0, maxiter=0):
        if reltol is None:
            reltol = float(np.sqrt(np.finfo(np.float64).eps))
        b = _as(b, np.float64).ravel()
        x = np.empty(self.n_dofs)
        it = C.c_int64(0)
        res = C.c_double(0)
        self._check(self._lib.onsas_pcg(self._h, b, x, precond, reltol, abstol, maxiter, C.byref(it), C.byref(res)))
        return x, it.value, res.value

    def spmv(self, x):
        x = _as(x, np.float64).ravel()
        y = np.empty(self.n_dofs)
        self._check(self._lib.onsas_spmv(self._h, x, y))
        return y

    def spmv_resident(self):
        self._check(self._lib.onsas_spmv_resident(self._h))

    # -- results
    def get_csr(self):
        n = C.c_int64(0)
        nnz = C.c_int64(0)
        self._check(self._lib.onsas_get_csr_size(self._h, C.byref(n), C.byref(nnz)))
        rowptr = np.empty(n.value + 1, np.int64)
        col = np.empty(nnz.value, np.int32)
        val = np.empty(nnz.value)
        self._check(self._lib.onsas_get_csr(self._h, rowptr, col, val))
        return rowptr, col, val

    def get_stress_strain(self, family: int = L.FAMILY_TET):
        ne = self.n_tets if family == L.FAMILY_TET else self.n_trusses
        s = np.empty((ne, 9))
        e = np.empty((ne, 9))
        self._check(self._lib.onsas_get_stress_strain(self._h, family, s.ravel(), e.ravel()))
        return s, e

    def table_stats(self) -> dict:
        out = np.zeros(8, np.int64)
        self._check(self._lib.onsas_get_table_stats(self._h, out))
        keys = ["n_slices", "padded_block_slots", "nnz_blocks", "tet_pairs", "truss_pairs", "max_pairs_per_slice",
                "k_bytes", "cg_grid"]
        return dict(zip(keys, (int(v) for v in out)))

    def cg_profile(self) -> dict:
        out = np.zeros(8, np.int64)
        self._check(self._lib.onsas_get_cg_profile(self._h, out))
        keys = ["update_p", "sync1", "spmv_dot", "sync2", "update_xr", "sync3", "reductions", "slowest_cta_spmv"]
        return dict(zip(keys, (int(v) for v in out[:8])))

    # -- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        st = L.lib().onsas_comm_unique_id(C.cast(buf, C.c_void_p))
        if st != L.OK:
            raise OnsasError(st, "ncclGetUniqueId failed")
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._lib.onsas_comm_init(self._h, n_ranks, rank, C.cast(buf, C.c_void_p)))

    def set_halo(self, nbr_rank, send_ptr, send_nodes, recv_ptr):
        nbr = _as(nbr_rank, np.int32)
        sp = _as(send_ptr, np.int64)
        sn = _as(send_nodes, np.int32)
        rp = _as(recv_ptr, np.int64)
        self._check(self._lib.onsas_set_halo(self._h, len(nbr), _ptr(nbr), _ptr(sp), _ptr(sn), _ptr(rp)))


    def p2p_export(self):
        buf = C.create_string_buffer(64)
        off = C.c_int64(0)
        self._check(self._lib.onsas_p2p_export(self._h, C.cast(buf, C.c_void_p), C.byref(off)))
        return buf.raw, off.value

    def p2p_import(self, handles, offsets, remote_halo_node_off=None):
        """`remote_halo_node_off` may be None for a context loaded from a NativePartition (it knows the offsets)."""
        blob = C.create_string_buffer(b"".join(handles), 64 * len(handles))
        offs = _as(offsets, np.int64)
        rem = None if remote_halo_node_off is None else _as(remote_halo_node_off, np.int64)
        self._check(self._lib.onsas_p2p_import(self._h, C.cast(blob, C.c_void_p), offs, _ptr(rem) if rem is not None and len(rem) else None))

    def load_part(self, part: "NativePartition", rank: int, truss_strain: int = 0):
        """Nodes, elements, free dofs and halo plan of `rank` from a native partition (onsas_part_load)."""
        self._check(self._lib.onsas_part_load(part._h, int(rank), self._h, int(truss_strain)))
        sz = part.sizes(rank)
        self.n_nodes, self.n_owned, self.n_tets, self.n_trusses = sz["n_local"], sz["n_owned"], sz["n_tets"], sz["n_trusses"]
        self.dim = part.dim


class NativePartition:
    """The library's partitioner (csrc/partition.cpp) on its own: recursive coordinate bisection of the global mesh into
    `n_ranks` parts, the halo plan of every rank.  Used by the one-process-per-GPU binding (every process builds the same
    partition and loads its own rank); a multi-device DeviceContext runs the same code inside onsas_finalize_mesh."""

    def __init__(self, xyz, n_ranks: int, tets=None, tet_mat=None, trusses=None, truss_mat=None, truss_area=None, free_dofs=None, reorder: int = 0):
        self._lib = L.lib()
        xyz = _as(xyz, np.float64)
        if xyz.ndim == 1:
            xyz = xyz.reshape(-1, 1)
        self.n_nodes, self.dim = xyz.shape
        self.n_ranks = int(n_ranks)
        tets = None if tets is None or len(tets) == 0 else _as(tets, np.int32).reshape(-1, 4)
        trusses = None if trusses is None or len(trusses) == 0 else _as(trusses, np.int32).reshape(-1, 2)
        tm = None if tet_mat is None else _as(tet_mat, np.int32)
        bm = None if truss_mat is None else _as(truss_mat, np.int32)
        ar = None if truss_area is None else _as(truss_area, np.float64)
        fd = np.arange(self.n_nodes * self.dim, dtype=np.int64) if free_dofs is None else _as(free_dofs, np.int64).ravel()
        h = C.c_void_p()
        st = self._lib.onsas_part_create(self.dim, self.n_nodes, xyz.ravel(), 0 if tets is None else len(tets), _ptr(tets), _ptr(tm),
                                         0 if trusses is None else len(trusses), _ptr(trusses), _ptr(bm), _ptr(ar), len(fd), _ptr(fd),
                                         self.n_ranks, int(reorder), C.byref(h))
        if st != L.OK:
            raise OnsasError(st, (self._lib.onsas_last_error(None) or b"").decode())
        self._h = h

    def _check(self, st):
        if st != L.OK:
            raise OnsasError(st, (self._lib.onsas_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.onsas_part_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sizes(self, rank: int) -> dict:
        out = np.zeros(8, np.int64)
        self._check(self._lib.onsas_part_sizes(self._h, int(rank), out))
        keys = ["n_local", "n_owned", "n_tets", "n_trusses", "n_free", "n_nbr", "n_send", "n_free_global"]
        return dict(zip(keys, (int(v) for v in out)))

    def local_to_global(self, rank: int) -> np.ndarray:
        l2g = np.empty(self.sizes(rank)["n_local"], np.int32)
        self._check(self._lib.onsas_part_local_to_global(self._h, int(rank), l2g))
        return l2g

    def local_elements(self, rank: int, family: int = L.FAMILY_TET) -> np.ndarray:
        sz = self.sizes(rank)
        gid = np.empty(sz["n_tets"] if family == L.FAMILY_TET else sz["n_trusses"], np.int64)
        self._check(self._lib.onsas_part_local_elements(self._h, int(rank), int(family), gid))
        return gid

    def halo_plan(self, rank: int) -> dict:
        sz = self.sizes(rank)
        nn = sz["n_nbr"]
        nbr, sp, sn = np.empty(nn, np.int32), np.empty(nn + 1, np.int64), np.empty(sz["n_send"], np.int32)
        rp, ro = np.empty(nn + 1, np.int64), np.empty(nn, np.int64)
        self._check(self._lib.onsas_part_halo_plan(self._h, int(rank), _ptr(nbr), _ptr(sp), _ptr(sn), _ptr(rp), _ptr(ro)))
        return dict(nbr_rank=nbr, send_ptr=sp, send_nodes=sn, recv_ptr=rp, remote_halo_off=ro)


def context_from_flat(xyz, tets=None, trusses=None, truss_area=None, truss_strain=0, mat_kind=(0,), mat_params=((1.0, 1.0),),
                      tet_mat=None, truss_mat=None, free_dofs=None, device=0, n_owned=None,
                      n_free_global: int = 0, reorder: int = 0) -> DeviceContext:
    """Upload a structure-of-arrays model and finalize it.  `reorder = 1`: the library renumbers the nodes along a Z-curve
    internally (ONSAS_OPT_REORDER; the caller keeps its own numbering everywhere)."""
    ctx = DeviceContext(device)
    if reorder:
        ctx.set_option(L.OPT_REORDER, reorder)
    ctx.set_nodes(xyz, n_owned)
    ctx.set_materials(mat_kind, mat_params)
    if tets is not None and len(tets):
        ctx.set_tets(tets, tet_mat)
    if trusses is not None and len(trusses):
        ctx.set_trusses(trusses, truss_area, truss_mat, truss_strain)
    if free_dofs is None:
        free_dofs = np.arange(ctx.n_owned * ctx.dim, dtype=np.int64)
    ctx.set_free_dofs(free_dofs, n_free_global)
    ctx.finalize()
    return ctx
