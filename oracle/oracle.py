"""ctypes binding + Newton driver for the CPU oracle (oracle/onsas_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of onsas_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.

The Newton driver restates the reference's control flow exactly
(StructuralAnalyses/NonLinearStaticAnalyses.jl:70-148, StructuralSolvers/StructuralSolvers.jl:115-171):
lagged residual test, ||U|| taken before the update, test order residual -> dU -> max_iter.
The linear solve is either scipy's direct sparse solve (the "direct-solve result" the
north star compares against) or the restated IterativeSolvers.jl CG in the C file.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

MAT_SVK, MAT_NEOHOOKEAN, MAT_ISOLINEAR = 0, 1, 2
STRAIN_ROTENG, STRAIN_GREEN = 0, 1
ORC_OK, ORC_ERR_NEG_VOLUME = 0, 1

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc -O2 -fopenmp)."""
    src = os.path.join(_HERE, "onsas_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_num_threads.restype = C.c_int
        L.orc_svk_stress.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_neohookean_stress.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_isolinear_stress.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp]
        L.orc_strain_energy.argtypes = [C.c_int, C.c_double, C.c_double, _dp]
        L.orc_strain_energy.restype = C.c_double
        L.orc_tet_volume.argtypes = [_dp]
        L.orc_tet_volume.restype = C.c_double
        L.orc_tet_internal_forces.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_truss_internal_forces.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_eval_tets.argtypes = [C.c_int64, _i32p, _i32p, _i32p, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_eval_trusses.argtypes = [C.c_int64, C.c_int, C.c_int, _i32p, _i32p, _i32p, _dp, _dp, _dp, _dp, _dp, _dp,
                                       _dp, _dp]
        L.orc_pattern_build.argtypes = [C.c_int64, C.c_int64, C.c_int, _i64p, _i64p, C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_assemble_family.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, _i32p, _i32p, _i32p, _dp, C.c_void_p,
                                          _dp, _dp, C.c_int64, _i64p, _i32p, _dp, _dp, _dp, _dp, C.c_int]
        L.orc_assemble_tets_mt.argtypes = [C.c_int64, _i32p, _i32p, _i32p, _dp, _dp, _dp, C.c_int64, _i64p, _i32p,
                                           _i64p, _i32p, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_cg.argtypes = [C.c_int64, _i64p, _i32p, _dp, _u8p, _dp, _dp, C.c_void_p, C.c_double, C.c_double,
                             C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        for fn in ("orc_tet_internal_forces", "orc_truss_internal_forces", "orc_eval_tets", "orc_eval_trusses",
                   "orc_pattern_build", "orc_assemble_family", "orc_assemble_tets_mt", "orc_cg"):
            getattr(L, fn).restype = C.c_int
        _lib = L
    return _lib


class NegativeVolumeError(ValueError):
    """Mirror of ArgumentError("Element with negative volume, check connectivity.") (Tetrahedrons.jl:136)."""


def _check(status: int):
    if status == ORC_ERR_NEG_VOLUME:
        raise NegativeVolumeError("Element with negative volume, check connectivity.")
    if status != ORC_OK:
        raise RuntimeError(f"oracle error status {status}")


# ----------------------------------------------------------------------------- single items

def material_stress(kind: int, p0: float, p1: float, E: np.ndarray):
    """(S, D) for a 3x3 strain E; D is 6x6 in the reference's Voigt order."""
    Ecm = np.asfortranarray(np.asarray(E, dtype=np.float64)).ravel(order="F").copy()
    S = np.zeros(9)
    D = np.zeros(36)
    fn = {MAT_SVK: lib().orc_svk_stress, MAT_NEOHOOKEAN: lib().orc_neohookean_stress,
          MAT_ISOLINEAR: lib().orc_isolinear_stress}[kind]
    fn(p0, p1, Ecm, S, D)
    return S.reshape(3, 3, order="F"), D.reshape(6, 6, order="F")


def strain_energy(kind: int, p0: float, p1: float, E: np.ndarray) -> float:
    Ecm = np.asarray(E, dtype=np.float64).ravel(order="F").copy()
    return lib().orc_strain_energy(kind, p0, p1, Ecm)


def tet_volume(X: np.ndarray) -> float:
    """X: 4x3 array of node coordinates."""
    return lib().orc_tet_volume(np.ascontiguousarray(X, dtype=np.float64).ravel())


def tet_internal_forces(kind, p0, p1, X, u):
    """X: (4,3) node coords; u: (12,) node-major.  Returns f(12), K(12,12), sig(3,3), eps(3,3)."""
    X = np.ascontiguousarray(X, dtype=np.float64).ravel()
    u = np.ascontiguousarray(u, dtype=np.float64).ravel()
    f = np.zeros(12)
    K = np.zeros(144)
    s = np.zeros(9)
    e = np.zeros(9)
    _check(lib().orc_tet_internal_forces(kind, p0, p1, X, u, f, K, s, e))
    return f, K.reshape(12, 12, order="F"), s.reshape(3, 3, order="F"), e.reshape(3, 3, order="F")


def truss_internal_forces(strain_model, dim, Emod, A, X, u):
    """X: (2,dim); u: (2*dim,)."""
    X = np.ascontiguousarray(X, dtype=np.float64).ravel()
    u = np.ascontiguousarray(u, dtype=np.float64).ravel()
    n = 2 * dim
    f = np.zeros(n)
    K = np.zeros(n * n)
    s = np.zeros(9)
    e = np.zeros(9)
    _check(lib().orc_truss_internal_forces(strain_model, dim, Emod, A, X, u, f, K, s, e))
    return f, K.reshape(n, n, order="F"), s.reshape(3, 3, order="F"), e.reshape(3, 3, order="F")


# ----------------------------------------------------------------------------- flat model

@dataclass
class FlatModel:
    """Structure-of-arrays restatement of a `Structure` (what the C ABI receives).

    xyz (n_nodes, dim); tets (n,4) int32; trusses (n,2) int32; materials: kind[], params (m,2).
    free_dofs: int64 0-based, node order with fixed removed (Structures.jl:129-142)."""
    xyz: np.ndarray
    dim: int = 3
    tets: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.int32))
    tet_mat: np.ndarray | None = None
    trusses: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    truss_mat: np.ndarray | None = None
    truss_area: np.ndarray | None = None
    truss_strain: int = STRAIN_ROTENG
    mat_kind: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    mat_params: np.ndarray = field(default_factory=lambda: np.zeros((1, 2)))
    free_dofs: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))

    def __post_init__(self):
        self.xyz = np.ascontiguousarray(self.xyz, dtype=np.float64).reshape(-1, self.dim)
        self.tets = np.ascontiguousarray(self.tets, dtype=np.int32).reshape(-1, 4)
        self.trusses = np.ascontiguousarray(self.trusses, dtype=np.int32).reshape(-1, 2)
        self.mat_kind = np.ascontiguousarray(self.mat_kind, dtype=np.int32)
        self.mat_params = np.ascontiguousarray(self.mat_params, dtype=np.float64).reshape(-1, 2)
        if self.tet_mat is None:
            self.tet_mat = np.zeros(len(self.tets), np.int32)
        if self.truss_mat is None:
            self.truss_mat = np.zeros(len(self.trusses), np.int32)
        if self.truss_area is None:
            self.truss_area = np.ones(len(self.trusses))
        self.tet_mat = np.ascontiguousarray(self.tet_mat, dtype=np.int32)
        self.truss_mat = np.ascontiguousarray(self.truss_mat, dtype=np.int32)
        self.truss_area = np.ascontiguousarray(self.truss_area, dtype=np.float64)
        self.free_dofs = np.ascontiguousarray(self.free_dofs, dtype=np.int64)

    @property
    def n_nodes(self):
        return self.xyz.shape[0]

    @property
    def n_dofs(self):
        return self.n_nodes * self.dim

    def free_mask(self):
        m = np.zeros(self.n_dofs, np.uint8)
        m[self.free_dofs] = 1
        return m


def eval_tets(m: FlatModel, U: np.ndarray):
    n = len(m.tets)
    f = np.zeros((n, 12))
    K = np.zeros((n, 144))
    s = np.zeros((n, 9))
    e = np.zeros((n, 9))
    _check(lib().orc_eval_tets(n, m.tets, m.tet_mat, m.mat_kind, m.mat_params.ravel(), m.xyz.ravel(),
                               np.ascontiguousarray(U, np.float64), f.ravel(), K.ravel(), s.ravel(), e.ravel()))
    return f, K, s, e


def eval_trusses(m: FlatModel, U: np.ndarray):
    n = len(m.trusses)
    nd = 2 * m.dim
    f = np.zeros((n, nd))
    K = np.zeros((n, nd * nd))
    s = np.zeros((n, 9))
    e = np.zeros((n, 9))
    _check(lib().orc_eval_trusses(n, m.dim, m.truss_strain, m.trusses, m.truss_mat, m.mat_kind, m.mat_params.ravel(),
                                  m.truss_area, m.xyz.ravel(), np.ascontiguousarray(U, np.float64), f.ravel(),
                                  K.ravel(), s.ravel(), e.ravel()))
    return f, K, s, e


class Assembly:
    """Reference-order serial assembly into a fixed CSR pattern (pattern = what the reference's
    first end_assemble! leaves in its SparseMatrixCSC)."""

    def __init__(self, m: FlatModel):
        self.m = m
        d = m.dim
        lists = []
        if len(m.tets):
            lists.append((m.tets.astype(np.int64)[:, :, None] * d + np.arange(d)[None, None, :]).reshape(len(m.tets), -1))
        if len(m.trusses):
            lists.append((m.trusses.astype(np.int64)[:, :, None] * d + np.arange(d)[None, None, :]).reshape(len(m.trusses), -1))
        n = m.n_dofs
        # union pattern over families (pattern_build handles one nde at a time -> merge via scipy-free numpy)
        rows, cols = [], []
        self.rowptr = np.zeros(n + 1, np.int64)
        if len(lists) == 1:
            ed = np.ascontiguousarray(lists[0])
            nnz = C.c_int64(0)
            _check(lib().orc_pattern_build(n, ed.shape[0], ed.shape[1], ed.ravel(), self.rowptr, None, C.byref(nnz)))
            self.col = np.zeros(nnz.value, np.int32)
            _check(lib().orc_pattern_build(n, ed.shape[0], ed.shape[1], ed.ravel(), self.rowptr,
                                           self.col.ctypes.data_as(C.c_void_p), C.byref(nnz)))
        else:
            for ed in lists:
                nde = ed.shape[1]
                rows.append(np.repeat(ed, nde, axis=1).ravel())
                cols.append(np.tile(ed, (1, nde)).ravel())
            r = np.concatenate(rows) if rows else np.zeros(0, np.int64)
            c = np.concatenate(cols) if cols else np.zeros(0, np.int64)
            key = np.unique(r * n + c)
            r, c = key // n, key % n
            self.rowptr[1:] = np.cumsum(np.bincount(r, minlength=n))
            self.col = c.astype(np.int32)
        self.val = np.zeros(len(self.col))
        self.F_int = np.zeros(n)
        self.tet_sig = np.zeros((len(m.tets), 9))
        self.tet_eps = np.zeros((len(m.tets), 9))
        self.truss_sig = np.zeros((len(m.trusses), 9))
        self.truss_eps = np.zeros((len(m.trusses), 9))

    def assemble(self, U: np.ndarray):
        """reset_assemble! + element loop + end_assemble! (StaticAnalyses.jl:99-132)."""
        m = self.m
        U = np.ascontiguousarray(U, np.float64)
        self.F_int[:] = 0
        self.val[:] = 0
        if len(m.tets):
            _check(lib().orc_assemble_family(0, len(m.tets), 3, 0, m.tets, m.tet_mat, m.mat_kind, m.mat_params.ravel(),
                                             None, m.xyz.ravel(), U, m.n_dofs, self.rowptr, self.col, self.val,
                                             self.F_int, self.tet_sig.ravel(), self.tet_eps.ravel(), 0))
        if len(m.trusses):
            _check(lib().orc_assemble_family(1, len(m.trusses), m.dim, m.truss_strain, m.trusses, m.truss_mat,
                                             m.mat_kind, m.mat_params.ravel(),
                                             m.truss_area.ctypes.data_as(C.c_void_p), m.xyz.ravel(), U, m.n_dofs,
                                             self.rowptr, self.col, self.val, self.F_int, self.truss_sig.ravel(),
                                             self.truss_eps.ravel(), 0))
        return self

    def csr(self):
        import scipy.sparse as sp
        n = self.m.n_dofs
        return sp.csr_matrix((self.val.copy(), self.col.copy(), self.rowptr.copy()), shape=(n, n))

    def dense(self):
        return self.csr().toarray()


class AssemblyMT:
    """All-host-threads tet assembly (OpenMP); same per-entry summation order as `Assembly`."""

    def __init__(self, m: FlatModel):
        self.base = Assembly(m)
        self.m = m
        ne = len(m.tets)
        pairs = (np.arange(ne, dtype=np.int64)[:, None] * 4 + np.arange(4)[None, :]).ravel()
        nodes = m.tets.astype(np.int64).ravel()
        order = np.argsort(nodes, kind="stable")
        self.adj = pairs[order].astype(np.int32)
        self.adj_ptr = np.zeros(m.n_nodes + 1, np.int64)
        self.adj_ptr[1:] = np.cumsum(np.bincount(nodes, minlength=m.n_nodes))
        self.scratchK = np.zeros(ne * 144)
        self.scratchf = np.zeros(ne * 12)

    def assemble(self, U):
        b, m = self.base, self.m
        _check(lib().orc_assemble_tets_mt(len(m.tets), m.tets, m.tet_mat, m.mat_kind, m.mat_params.ravel(),
                                          m.xyz.ravel(), np.ascontiguousarray(U, np.float64), m.n_nodes, self.adj_ptr,
                                          self.adj, b.rowptr, b.col, b.val, b.F_int, b.tet_sig.ravel(),
                                          b.tet_eps.ravel(), self.scratchK, self.scratchf))
        return b


def cg(rowptr, col, val, free_mask, b, diag=None, reltol=None, abstol=0.0, maxiter=None):
    """IterativeSolvers.jl cg restatement; defaults = StructuralSolvers.jl:229-234."""
    n = len(b)
    if reltol is None:
        reltol = float(np.sqrt(np.finfo(np.float64).eps))
    if maxiter is None:
        maxiter = int(np.count_nonzero(free_mask))
    x = np.zeros(n)
    it = C.c_int64(0)
    res = C.c_double(0)
    dptr = None if diag is None else np.ascontiguousarray(diag, np.float64).ctypes.data_as(C.c_void_p)
    _check(lib().orc_cg(n, rowptr, col, np.ascontiguousarray(val), np.ascontiguousarray(free_mask, np.uint8),
                        np.ascontiguousarray(b, np.float64), x, dptr, reltol, abstol, maxiter, C.byref(it),
                        C.byref(res)))
    return x, it.value, res.value


# ----------------------------------------------------------------------------- Newton driver

@dataclass
class ConvergenceSettings:  # StructuralSolvers.jl:36-43 (positional order U, force, iter)
    rel_U_tol: float = 1e-6
    rel_res_force_tol: float = 1e-6
    max_iter: int = 20


INITIAL_DELTA = 1e12  # StructuralSolvers.jl:27


def criterion(dU_rel, dr_rel, it, tols: ConvergenceSettings) -> str:
    """isconverged! (StructuralSolvers.jl:148-171)."""
    assert dU_rel > 0, "Residual displacements norm must be greater than 0."
    assert dr_rel > 0, "Residual forces norm must be greater than 0."
    if dr_rel <= tols.rel_res_force_tol:
        return "ResidualForceCriterion"
    if dU_rel <= tols.rel_U_tol:
        return "DeltaUCriterion"
    if it > tols.max_iter:
        return "MaxIterCriterion"
    return "NotConvergedYet"


@dataclass
class NewtonResult:
    U: list            # per load step: full displacement vector
    F_int: list        # per load step: internal forces at the last assembly (reactions at fixed dofs)
    iterations: list   # Newton iterations per load step
    criteria: list
    cg_iterations: list
    tet_sig: list
    tet_eps: list
    truss_sig: list
    truss_eps: list
    norms: list        # last (dU_norm, dU_rel, r_norm, r_rel) per load step


def newton_solve(m: FlatModel, load_factors, fext_fn, tols: ConvergenceSettings, linear="direct",
                 cg_reltol=None, jacobi=False, U0=None, assembler=None) -> NewtonResult:
    """_solve!(::NonLinearStaticAnalysis, ::NewtonRaphson, ...) (NonLinearStaticAnalyses.jl:70-104) and
    step! (:107-148).  fext_fn(t) -> full F_ext vector (apply!, StructuralAnalyses.jl:228-241)."""
    import scipy.sparse.linalg as spla
    asm = assembler or Assembly(m)
    base = asm.base if isinstance(asm, AssemblyMT) else asm
    n = m.n_dofs
    free = m.free_dofs
    mask = m.free_mask()
    U = np.zeros(n) if U0 is None else np.array(U0, dtype=np.float64)
    out = NewtonResult([], [], [], [], [], [], [], [], [], [])
    for t in load_factors:
        dU_norm = dr_norm = dU_rel = dr_rel = INITIAL_DELTA  # reset! (:115-121)
        it = 0
        F_ext = np.asarray(fext_fn(t), dtype=np.float64)
        cg_its = []
        crit = criterion(dU_rel, dr_rel, it, tols)
        while crit == "NotConvergedYet":
            asm.assemble(U)
            r = F_ext[free] - base.F_int[free]  # residual_forces! (StaticStates.jl:113-116)
            if linear == "direct":
                A = base.csr()[free][:, free].tocsc()
                dU = spla.spsolve(A, r)
            else:
                b = np.zeros(n)
                b[free] = r
                diag = None
                if jacobi:
                    diag = base.csr().diagonal().copy()
                    diag[mask == 0] = 1.0
                x, its, _ = cg(base.rowptr, base.col, base.val, mask, b, diag=diag, reltol=cg_reltol)
                cg_its.append(its)
                dU = x[free]
            dU_norm = float(np.linalg.norm(dU))
            nU = float(np.linalg.norm(U))
            with np.errstate(divide="ignore"):
                dU_rel = dU_norm / nU if nU > 0 else float("inf")  # Julia: x/0.0 = Inf
            dr_norm = float(np.linalg.norm(r))
            dr_rel = dr_norm / float(np.linalg.norm(F_ext))
            U[free] += dU
            it += 1
            crit = criterion(dU_rel, dr_rel, it, tols)
        out.U.append(U.copy())
        out.F_int.append(base.F_int.copy())
        out.iterations.append(it)
        out.criteria.append(crit)
        out.cg_iterations.append(cg_its)
        out.tet_sig.append(base.tet_sig.copy())
        out.tet_eps.append(base.tet_eps.copy())
        out.truss_sig.append(base.truss_sig.copy())
        out.truss_eps.append(base.truss_eps.copy())
        out.norms.append((dU_norm, dU_rel, dr_norm, dr_rel))
    return out
