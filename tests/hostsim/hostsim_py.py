"""ctypes wrapper of the TEST-ONLY host simulation (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os

import numpy as np

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhostsim.so")
_dp = np.ctypeslib.ndpointer(np.float64, flags="C")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C")
_lp = np.ctypeslib.ndpointer(np.int64, flags="C")
_u8 = np.ctypeslib.ndpointer(np.uint8, flags="C")
_vp = C.c_void_p

hs = C.CDLL(_SO)
hs.hs_create.restype = _vp
hs.hs_create.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int64, _vp]
hs.hs_error.restype = C.c_char_p
hs.hs_error.argtypes = [_vp]
hs.hs_destroy.argtypes = [_vp]
hs.hs_nnz.restype = C.c_int64
hs.hs_tables_hash.argtypes = [_vp]
hs.hs_tables_hash.restype = C.c_uint64
hs.hs_nnz.argtypes = [_vp]
hs.hs_stat.restype = C.c_int64
hs.hs_stat.argtypes = [_vp, C.c_int]
hs.hs_host_plan.restype = C.c_int
hs.hs_check_slice_nodes.restype = None
hs.hs_check_slice_nodes.argtypes = [_vp, C.c_int, _vp, C.c_int64, _lp]
hs.hs_host_plan.argtypes = [_vp, C.c_int, C.c_int, _lp, _lp]
hs.hs_assemble.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _ip, _dp, _dp, _dp, C.c_int]
hs.hs_get.argtypes = [_vp, _lp, _ip, _dp, _dp, _dp, _dp]
hs.hs_spmv.argtypes = [_vp, _u8, _dp, _dp]
hs.hs_pcg.argtypes = [_vp, _u8, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_int64),
                      C.POINTER(C.c_double)]
hs.hs_eval_tets.argtypes = [C.c_int64, _ip, _vp, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
hs.hs_eval_trusses.argtypes = [C.c_int64, C.c_int, C.c_int, _ip, _vp, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]


def _p(a):
    return None if a is None or len(a) == 0 else a.ctypes.data_as(_vp)


class HostSim:
    """Walks the device algorithms on the CPU for a flat model (oracle.FlatModel)."""

    def __init__(self, m, n_rows=None):
        self.m = m
        self.n_rows = m.n_nodes if n_rows is None else n_rows
        self.h = hs.hs_create(m.dim, m.n_nodes, self.n_rows, len(m.tets), _p(m.tets), len(m.trusses), _p(m.trusses))
        err = hs.hs_error(self.h).decode()
        if err:
            raise ValueError(err)

    def __del__(self):
        if getattr(self, "h", None):
            hs.hs_destroy(self.h)
            self.h = None

    def tables_hash(self):
        return int(hs.hs_tables_hash(self.h))

    def stats(self):
        return [hs.hs_stat(self.h, i) for i in range(5)]

    def check_slice_nodes(self, family, conn):
        out = np.zeros(6, np.int64)
        conn = np.ascontiguousarray(conn, np.int32)
        hs.hs_check_slice_nodes(self.h, family, conn.ctypes.data_as(_vp), len(conn), out)
        return dict(zip(("violations", "unsorted", "bad_index", "bad_writer", "max_list", "entries"), (int(v) for v in out)))

    def host_plan(self, chunks, mid_weight):
        """(slice0 [n+1], node_hi [n]) of onsas_assemble_host's slice ranges."""
        s0, hi = np.zeros(chunks + 1, np.int64), np.zeros(chunks, np.int64)
        n = hs.hs_host_plan(self.h, chunks, mid_weight, s0, hi)
        return s0[:n + 1], hi[:n]

    def assemble(self, U, threads=192):
        m = self.m
        hs.hs_assemble(self.h, _p(m.tets), _p(m.tet_mat), _p(m.trusses), _p(m.truss_mat), _p(m.truss_area),
                       m.truss_strain, m.mat_kind, m.mat_params.ravel(), m.xyz.ravel(),
                       np.ascontiguousarray(U, np.float64), threads)
        nnz = hs.hs_nnz(self.h)
        nr = self.n_rows * m.dim
        rp = np.zeros(nr + 1, np.int64)
        ci = np.zeros(nnz, np.int32)
        v = np.zeros(nnz)
        Fi = np.zeros(nr)
        to = np.zeros(max(len(m.tets) * 16, 1))
        tr = np.zeros(max(len(m.trusses) * 2, 1))
        hs.hs_get(self.h, rp, ci, v, Fi, to, tr)
        return rp, ci, v, Fi, to[:len(m.tets) * 16].reshape(-1, 16), tr[:len(m.trusses) * 2].reshape(-1, 2)

    def spmv(self, mask, x):
        y = np.zeros(self.n_rows * self.m.dim)
        hs.hs_spmv(self.h, np.ascontiguousarray(mask, np.uint8), np.ascontiguousarray(x, np.float64), y)
        return y

    def pcg(self, mask, b, precond, reltol, abstol=0.0, maxiter=None):
        x = np.zeros(self.n_rows * self.m.dim)
        it = C.c_int64()
        res = C.c_double()
        if maxiter is None:
            maxiter = int(np.count_nonzero(mask))
        hs.hs_pcg(self.h, np.ascontiguousarray(mask, np.uint8), np.ascontiguousarray(b, np.float64), x, precond, reltol,
                  abstol, maxiter, C.byref(it), C.byref(res))
        return x, it.value, res.value


def eval_tets(m, U):
    n = len(m.tets)
    f, K, s, e = np.zeros((n, 12)), np.zeros((n, 144)), np.zeros((n, 9)), np.zeros((n, 9))
    hs.hs_eval_tets(n, m.tets, _p(m.tet_mat), m.mat_kind, m.mat_params.ravel(), m.xyz.ravel(),
                    np.ascontiguousarray(U, np.float64), f.ravel(), K.ravel(), s.ravel(), e.ravel())
    return f, K, s, e


def eval_trusses(m, U):
    n = len(m.trusses)
    nd = 2 * m.dim
    f, K, s, e = np.zeros((n, nd)), np.zeros((n, nd * nd)), np.zeros((n, 9)), np.zeros((n, 9))
    hs.hs_eval_trusses(n, m.dim, m.truss_strain, m.trusses, _p(m.truss_mat), m.mat_kind, m.mat_params.ravel(),
                       m.truss_area, m.xyz.ravel(), np.ascontiguousarray(U, np.float64), f.ravel(), K.ravel(), s.ravel(),
                       e.ravel())
    return f, K, s, e
