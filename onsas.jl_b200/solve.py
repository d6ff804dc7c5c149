"""Host-side mirror of the reference's solver layer for the hot path (same names, argument
meaning, control flow and error behaviour): ConvergenceSettings / ResidualsIterationStep /
isconverged! (StructuralSolvers/StructuralSolvers.jl:36-171), NewtonRaphson (Solvers.jl:14-29),
NonLinearStaticAnalysis and its `_solve!` / `step!` (StructuralAnalyses/NonLinearStaticAnalyses.jl:32-148),
LinearStaticAnalysis (LinearStaticAnalyses.jl:78-153), Solution accessors (Solutions.jl).

The loops stay on the host exactly as in the reference; each Newton iteration is ONE call into
libonsas_cuda (`onsas_newton_step` = assemble! + step!).  The state lives on the GPU for the whole solve.
"""
from __future__ import annotations

import copy
import math
import warnings
from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .device import DeviceContext
from .model import Structure

INITIAL_DELTA = 1e12  # StructuralSolvers.jl:27


@dataclass
class ConvergenceSettings:
    """Positional order is (rel_U_tol, rel_res_force_tol, max_iter) (StructuralSolvers.jl:36-43)."""
    rel_U_tol: float = 1e-6
    rel_res_force_tol: float = 1e-6
    max_iter: int = 20


class AbstractConvergenceCriterion:
    def __repr__(self):
        return type(self).__name__ + "()"

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)


class ResidualForceCriterion(AbstractConvergenceCriterion):
    pass


class ΔUCriterion(AbstractConvergenceCriterion):
    pass


DeltaUCriterion = ΔUCriterion


class MaxIterCriterion(AbstractConvergenceCriterion):
    pass


class NotConvergedYet(AbstractConvergenceCriterion):
    pass


class ResidualsIterationStep:
    """StructuralSolvers.jl:85-145."""

    def __init__(self):
        self.reset()

    def reset(self):  # :115-121
        self.ΔU_norm = self.Δr_norm = self.ΔU_rel = self.Δr_rel = INITIAL_DELTA
        self.iter = 0
        self.criterion = NotConvergedYet()
        return self

    def update(self, ΔU_norm, ΔU_rel, Δr_norm, Δr_rel):  # :136-145
        self.ΔU_norm, self.ΔU_rel, self.Δr_norm, self.Δr_rel = ΔU_norm, ΔU_rel, Δr_norm, Δr_rel
        self.iter += 1
        return self


def isconverged(ri: ResidualsIterationStep, cs: ConvergenceSettings):
    """isconverged! (StructuralSolvers.jl:148-171): test order residual -> dU -> max_iter."""
    assert ri.ΔU_rel > 0, "Residual displacements norm must be greater than 0."
    assert ri.Δr_rel > 0, "Residual forces norm must be greater than 0."
    if ri.Δr_rel <= cs.rel_res_force_tol:
        crit = ResidualForceCriterion()
    elif ri.ΔU_rel <= cs.rel_U_tol:
        crit = ΔUCriterion()
    elif ri.iter > cs.max_iter:
        warnings.warn("Maximum number of iterations was reached.")
        crit = MaxIterCriterion()
    else:
        crit = NotConvergedYet()
    ri.criterion = crit
    return crit


class NewtonRaphson:
    """NewtonRaphson(tols) (Solvers.jl:14-29) plus the knobs of the device linear solve:
    preconditioner "jacobi" (north star) or "none" (the reference's IterativeSolversJL_CG default),
    cg_reltol (default sqrt(eps), StructuralSolvers.jl:229-234), cg_abstol, cg_maxiter (0 = n_free)."""

    def __init__(self, tol: ConvergenceSettings | None = None, *, preconditioner="jacobi", cg_reltol=None,
                 cg_abstol=0.0, cg_maxiter=0, device=0, cg_mode=None):
        self.tol = tol or ConvergenceSettings()
        assert preconditioner in ("jacobi", "none", "two_level")
        self.preconditioner = preconditioner
        self.cg_reltol = math.sqrt(np.finfo(np.float64).eps) if cg_reltol is None else float(cg_reltol)
        self.cg_abstol, self.cg_maxiter, self.device, self.cg_mode = float(cg_abstol), int(cg_maxiter), device, cg_mode

    @property
    def precond_code(self):
        return {"jacobi": L.PRECOND_JACOBI, "two_level": L.PRECOND_TWO_LEVEL}.get(self.preconditioner, L.PRECOND_NONE)


NewtonRaphsonCUDA = NewtonRaphson


def tolerances(alg):
    return alg.tol


# ------------------------------------------------------------------------------------------ analyses

class _StaticAnalysisBase:
    def __init__(self, s: Structure, load_factors, initial_step=1):
        self.s = s
        self.λᵥ = np.asarray(load_factors, dtype=np.float64)
        if not (1 <= initial_step <= len(self.λᵥ)):
            raise ValueError(f"initial_step must be in [1, {len(self.λᵥ)}] but is: {initial_step}.")
        self.current_step = initial_step
        self.iter_state = ResidualsIterationStep()
        self._ctx: DeviceContext | None = None

    # StaticAnalyses.jl:60-96
    def structure(self):
        return self.s

    def load_factors(self):
        return self.λᵥ

    def current_time(self):
        return float(self.λᵥ[self.current_step - 1])

    def is_done(self):
        if self.current_step > len(self.λᵥ):
            self.current_step -= 1
            return True
        return False

    def next(self):
        self.current_step += 1

    def __deepcopy__(self, memo):
        # solve() deep-copies the analysis (StructuralSolvers.jl:201-206): never copy the raw device handle
        new = copy.copy(self)
        new.iter_state = copy.deepcopy(self.iter_state, memo)
        new.λᵥ = self.λᵥ.copy()
        new._ctx = None
        return new

    def device_context(self, device=0) -> DeviceContext:
        """Created lazily inside _solve! and uploaded once (`solve` deep-copy caveat)."""
        if self._ctx is None:
            f = self.s.flat
            ctx = DeviceContext(device)
            ctx.set_nodes(f.xyz)
            ctx.set_materials(f.mat_kind, f.mat_params)
            if len(f.tets):
                ctx.set_tets(f.tets, f.tet_mat)
            if len(f.trusses):
                ctx.set_trusses(f.trusses, f.truss_area, f.truss_mat, f.truss_strain)
            ctx.set_free_dofs(f.free_dofs)
            ctx.finalize()
            self._ctx = ctx
        return self._ctx


class NonLinearStaticAnalysis(_StaticAnalysisBase):
    """NonLinearStaticAnalysis(s, t1=1.0; NSTEPS=10) or (s, load_factors) (NonLinearStaticAnalyses.jl:32-60)."""

    def __init__(self, s: Structure, t1=1.0, *, NSTEPS=10, initial_step=1):
        if np.ndim(t1) > 0:
            lf = np.asarray(t1, dtype=np.float64)
        else:
            lf = np.linspace(t1 / NSTEPS, t1, NSTEPS)  # collect(LinRange(t0, t1, NSTEPS)) :55-60
        super().__init__(s, lf, initial_step)


class LinearStaticAnalysis(_StaticAnalysisBase):
    """LinearStaticAnalysis(s, t1=1.0; NSTEPS=10) (LinearStaticAnalyses.jl:32-60)."""

    def __init__(self, s: Structure, t1=1.0, *, NSTEPS=10, initial_step=1):
        lf = np.asarray(t1, dtype=np.float64) if np.ndim(t1) > 0 else np.linspace(t1 / NSTEPS, t1, NSTEPS)
        super().__init__(s, lf, initial_step)


# ------------------------------------------------------------------------------------------ solution

class Solution:
    """Per-load-step stored states (U, stress, strain) + accessors (Solutions.jl:58-91,139-185).
    Flat arrays instead of the reference's Dictionary-per-element."""

    def __init__(self, analysis, solver):
        self.analysis, self.solver = analysis, solver
        self.U: list[np.ndarray] = []
        self.F_int: list[np.ndarray] = []
        self.tet_stress: list[np.ndarray] = []
        self.tet_strain: list[np.ndarray] = []
        self.truss_stress: list[np.ndarray] = []
        self.truss_strain: list[np.ndarray] = []
        self._iterations: list[int] = []
        self._criteria: list = []
        self.cg_iterations: list[list[int]] = []
        self.step_info: list[list] = []

    def iterations(self):
        return list(self._iterations)

    def criterion(self):
        return list(self._criteria)

    def displacements(self, node=None, component=None):
        """displacements(sol) / (sol, node) / (sol, node, component) with a 1-based component."""
        if node is None:
            return self.U
        s = self.analysis.s
        i = s.node_index(node) if not isinstance(node, (int, np.integer)) else int(node)
        d = s.flat.dim
        per = [np.array([U[d * i + c] for U in self.U]) for c in range(d)]
        return per if component is None else per[component - 1]

    def reactions(self):
        """F_int at the fixed dofs of each stored step (the reference keeps them in F_int, StaticStates.jl:84-88)."""
        mask = np.ones(self.analysis.s.flat.n_dofs, dtype=bool)
        mask[self.analysis.s.flat.free_dofs] = False
        return [np.where(mask, F, 0.0) for F in self.F_int]

    def _elem(self, e, tets, trusses):
        fam, k = self.analysis.s.element_slot(e) if not isinstance(e, tuple) else e
        src = tets if fam == L.FAMILY_TET else trusses
        return [a[k].reshape(3, 3, order="F") for a in src]

    def stress(self, e):
        return self._elem(e, self.tet_stress, self.truss_stress)

    def strain(self, e):
        return self._elem(e, self.tet_strain, self.truss_strain)


# ------------------------------------------------------------------------------------------ solve

def solve(problem, solver=None, *args, **kwargs):
    """solve(problem, solver) = solve!(deepcopy(problem), solver) (StructuralSolvers.jl:201-206)."""
    return solve_(copy.deepcopy(problem), solver, *args, **kwargs)


def solve_(problem, solver=None, linear_solve=None, *, linear_solve_inplace=False):
    """solve!(problem, solver[, linear_solve]; linear_solve_inplace) (StructuralSolvers.jl:217-222).
    `linear_solve` is accepted for signature compatibility; the linear solve on this path is always
    the device (P)CG configured on the solver object."""
    if isinstance(problem, NonLinearStaticAnalysis):
        return _solve_nonlinear(problem, solver or NewtonRaphson())
    if isinstance(problem, LinearStaticAnalysis):
        return _solve_linear(problem, solver or NewtonRaphson())
    raise TypeError("unsupported analysis type")


def _store(sol: Solution, ctx: DeviceContext, sa):
    """store!(sol, state, step) (StaticAnalyses.jl:157-174) with flat arrays."""
    sol.U.append(ctx.get_U())
    sol.F_int.append(ctx.get_Fint())
    if ctx.n_tets:
        s, e = ctx.get_stress_strain(L.FAMILY_TET)
        sol.tet_stress.append(s)
        sol.tet_strain.append(e)
    if ctx.n_trusses:
        s, e = ctx.get_stress_strain(L.FAMILY_TRUSS)
        sol.truss_stress.append(s)
        sol.truss_strain.append(e)


def _solve_nonlinear(sa: NonLinearStaticAnalysis, alg: NewtonRaphson) -> Solution:
    """_solve!(::NonLinearStaticAnalysis, ...) (NonLinearStaticAnalyses.jl:70-104)."""
    s = sa.s
    ctx = sa.device_context(alg.device)
    if alg.cg_mode is not None:
        ctx.set_option(L.OPT_CG_MODE, alg.cg_mode)
    sol = Solution(sa, alg)
    while not sa.is_done():
        it = sa.iter_state.reset()                       # :83
        s.flat.apply_loads(ctx, sa.current_time())       # :86-87 external forces of this load step (built on the device)
        cg_its, infos = [], []
        while isinstance(isconverged(it, alg.tol), NotConvergedYet):   # :90
            info = ctx.newton_step(alg.precond_code, alg.cg_reltol, alg.cg_abstol, alg.cg_maxiter)  # :92-95
            rel_dU = info.norm_dU / info.norm_U if info.norm_U > 0 else math.inf   # :139 (x/0.0 = Inf in Julia)
            rel_r = info.norm_r / info.norm_Fext if info.norm_Fext > 0 else math.inf  # :141
            it.update(info.norm_dU, rel_dU, info.norm_r, rel_r)    # :147
            cg_its.append(int(info.cg_iters))
            infos.append((info.norm_dU, rel_dU, info.norm_r, rel_r, info.ms_assemble, info.ms_solve))
        _store(sol, ctx, sa)                             # :98
        sol._iterations.append(it.iter)
        sol._criteria.append(it.criterion)
        sol.cg_iterations.append(cg_its)
        sol.step_info.append(infos)
        sa.next()                                        # :101
    return sol


def _solve_linear(sa: LinearStaticAnalysis, alg: NewtonRaphson) -> Solution:
    """_solve!(::LinearStaticAnalysis, ...) (LinearStaticAnalyses.jl:78-153) in the reference's own sequence: K is assembled
    at the FIRST load step only (:93-96); every step solves K dU = F_ext(t)[free] and SETS U[free] = dU (step!, :117-153:
    onsas_step with update_U = 2), then the elements are re-evaluated at U for stress / strain (:101-102).  The reference
    solves with the copy of K it froze at step 1; the device holds one K, so when K depends on U (a hyperelastic material in
    a linear analysis) it is re-evaluated at the initial state before each later solve -- the same matrix, hence the same
    result; for IsotropicLinearElastic K does not depend on U and nothing is re-assembled."""
    s = sa.s
    ctx = sa.device_context(alg.device)
    sol = Solution(sa, alg)
    U0 = ctx.get_U()
    k_is_constant = all(int(k) == L.MAT_ISOLINEAR for k in s.flat.mat_kind) and len(s.flat.trusses) == 0
    first = True
    while not sa.is_done():
        s.flat.apply_loads(ctx, sa.current_time())       # :89-90
        if first or not k_is_constant:
            if not first:
                ctx.set_U(U0)
            ctx.assemble()                               # :93-96 K at the initial state
        info = ctx.step(alg.precond_code, alg.cg_reltol, alg.cg_abstol, alg.cg_maxiter, update_U=2)   # :99, :117-153
        ctx.assemble()                                   # stress / strain / F_int at the solved U (:101-102)
        ctx.synchronize()
        _store(sol, ctx, sa)
        sol._iterations.append(1)
        sol._criteria.append(ResidualForceCriterion())   # LinearResidualsIterationStep reset! :128-132
        sol.cg_iterations.append([int(info.cg_iters)])
        sa.next()
        first = False
    return sol
