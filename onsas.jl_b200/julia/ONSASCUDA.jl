# ONSASCUDA.jl -- Julia glue that makes libonsas_cuda a drop-in for ONSAS.jl's Newton-Raphson hot path.
#
# UNVERIFIED: no Julia toolchain exists in the build image or on the GPU box, so this file has never
# been executed.  It is written against ONSAS.jl v0.4.6 and the C ABI of include/onsas_cuda.h; the
# same ABI is exercised by the Python mirror (onsas.jl_b200/solve.py), which follows this file line by line.
#
# Where it goes in an ONSAS.jl checkout (north star: "Julia host code in src/StructuralSolvers and
# src/StructuralAnalyses calls a thin C-ABI shim"):
#   src/StructuralSolvers/CUDASolvers.jl   <- this file
#   src/ONSAS.jl                           <- add "StructuralSolvers/CUDASolvers.jl" to FILES after
#                                             "StructuralAnalyses/NonLinearStaticAnalyses.jl" and `@reexport using .CUDASolvers`
# User code changes one constructor:  NewtonRaphson(tols)  ->  NewtonRaphsonCUDA(tols)
# Everything else (Structure, NonLinearStaticAnalysis, solve / solve!, Solution accessors, write_vtk) is untouched.
module CUDASolvers

using LinearAlgebra, Dictionaries
using ..Utils, ..Nodes, ..Entities, ..Tetrahedrons, ..Trusses, ..CrossSections
using ..Materials, ..SVKMaterial, ..NeoHookeanMaterial, ..IsotropicLinearElasticMaterial
using ..Meshes, ..Structures, ..StructuralBoundaryConditions
using ..StructuralSolvers, ..Solvers, ..Solutions
using ..StructuralAnalyses, ..StaticStates, ..StaticAnalyses, ..NonLinearStaticAnalyses, ..LinearStaticAnalyses
using ..BoundaryConditions, ..LocalLoadBoundaryConditions, ..TriangularFaces

import ..StructuralSolvers: _solve!, tolerances

export NewtonRaphsonCUDA, libonsas_cuda_path!

const LIB = Ref{String}(get(ENV, "ONSAS_CUDA_LIB", "libonsas_cuda"))
libonsas_cuda_path!(p::AbstractString) = (LIB[] = String(p))

# status codes of include/onsas_cuda.h
const ONSAS_OK = Int32(0)
const ONSAS_ERR_NEGATIVE_VOLUME = Int32(2)

"Device-resident analogue of FullStaticState (StaticStates.jl:33-102): an opaque onsas_ctx*."
mutable struct CudaContext
    handle::Ptr{Cvoid}
    function CudaContext(device::Union{Integer, AbstractVector{<:Integer}} = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        st = if device isa Integer
            ccall((:onsas_create, LIB[]), Int32, (Int32, Ref{Ptr{Cvoid}}), device, h)
        else
            # one Julia process drives all the listed devices: global mesh and global vectors in, the partitioning,
            # the halo plan and the peer-memory wiring happen inside onsas_finalize_mesh (onsas_create_multi)
            devs = collect(Int32, device)
            GC.@preserve devs ccall((:onsas_create_multi, LIB[]), Int32, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devs, length(devs), h)
        end
        st == ONSAS_OK || error(unsafe_string(ccall((:onsas_last_error, LIB[]), Cstring, (Ptr{Cvoid},), C_NULL)))
        ctx = new(h[])
        finalizer(c -> (c.handle != C_NULL && ccall((:onsas_destroy, LIB[]), Int32, (Ptr{Cvoid},), c.handle);
                        c.handle = C_NULL), ctx)
    end
end
# `solve` deep-copies the analysis (StructuralSolvers.jl:201-206): a raw handle must never be duplicated.
Base.deepcopy_internal(c::CudaContext, ::IdDict) = c

function check(ctx::CudaContext, st::Int32)
    st == ONSAS_OK && return nothing
    msg = unsafe_string(ccall((:onsas_last_error, LIB[]), Cstring, (Ptr{Cvoid},), ctx.handle))
    # same error type and text as Tetrahedrons.jl:136
    st == ONSAS_ERR_NEGATIVE_VOLUME && throw(ArgumentError("Element with negative volume, check connectivity."))
    error("libonsas_cuda: $msg")
end

"""
Newton-Raphson on the GPU.  Same role as `NewtonRaphson` (Solvers.jl:14-29); extra fields configure the
device linear solve that replaces `IterativeSolversJL_CG` (StructuralSolvers.jl:29, NonLinearStaticAnalyses.jl:129-134).
"""
Base.@kwdef struct NewtonRaphsonCUDA <: AbstractSolver
    tol::ConvergenceSettings = ConvergenceSettings()
    preconditioner::Symbol = :jacobi    # :none = un-preconditioned CG (the reference default), :jacobi, :two_level
                                        # (Jacobi + aggregated coarse space, ONSAS_PRECOND_TWO_LEVEL)
    cg_reltol::Float64 = sqrt(eps())    # StructuralSolvers.jl:229-234
    cg_abstol::Float64 = 0.0
    cg_maxiter::Int = 0                 # 0 -> number of free dofs
    device::Union{Int, Vector{Int}} = 0 # one CUDA device, or several: `device = collect(0:7)` runs the same solve on 8 GPUs
end
NewtonRaphsonCUDA(tol::ConvergenceSettings; kw...) = NewtonRaphsonCUDA(; tol, kw...)
precond_code(alg::NewtonRaphsonCUDA) = Int32(alg.preconditioner === :none ? 0 : alg.preconditioner === :two_level ? 2 : 1)

"""
`assemble!(s, sa)` (StaticAnalyses.jl:99-122) on the device with the reference's host state on both sides:
`displacements(state)` goes in, `internal_forces(state)` comes back, K / stress / strain stay on the device
(fetch them with `onsas_get_csr` / `onsas_get_stress_strain` when needed).  One ccall; the two copies are pipelined
with the assembly kernel inside the library (onsas_assemble_host).
"""
function assemble_device!(ctx::CudaContext, state::FullStaticState)
    U, F = displacements(state), internal_forces(state)
    GC.@preserve U F check(ctx, ccall((:onsas_assemble_host, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        ctx.handle, U, F))
    state
end

material_code(m::SVK) = (Int32(0), collect(Float64, lame_parameters(m)))
material_code(m::NeoHookean) = (Int32(1), [bulk_modulus(m), shear_modulus(m)])
material_code(m::IsotropicLinearElastic) = (Int32(2), [elasticity_modulus(m), poisson_ratio(m)])
strain_code(::Type{RotatedEngineeringStrain}) = Int32(0)
strain_code(::Type{GreenStrain}) = Int32(1)

"Flatten a `Structure` once into the structure-of-arrays form of the C ABI and upload it."
function upload!(ctx::CudaContext, s::AbstractStructure)
    vnodes = nodes(s)
    dim = length(coordinates(first(vnodes)))
    n_nodes = length(vnodes)
    node_index = IdDict(n => i - 1 for (i, n) in enumerate(vnodes))          # 0-based
    xyz = Matrix{Float64}(undef, dim, n_nodes)
    for (i, n) in enumerate(vnodes)
        xyz[:, i] .= coordinates(n)
        # the library assumes the set_dofs!(mesh, :u, dim) numbering (Meshes.jl:85-98)
        dofs(n, :u) == collect(dim * (i - 1) .+ (1:dim)) ||
            error("libonsas_cuda needs dof = dim*(i-1)+c numbering for field :u")
    end
    kinds = Int32[]; params = Float64[]
    tets = Int32[]; tet_mat = Int32[]; tet_elems = Any[]
    bars = Int32[]; bar_mat = Int32[]; areas = Float64[]; bar_elems = Any[]
    strain = Int32(0)
    for (mi, (mat, elems)) in enumerate(pairs(materials(s)))                  # assembly order, StaticAnalyses.jl:105-106
        k, p = material_code(mat)
        push!(kinds, k); append!(params, p)
        for e in elems
            ids = Int32[node_index[n] for n in nodes(e)]
            if e isa Tetrahedron
                append!(tets, ids); push!(tet_mat, mi - 1); push!(tet_elems, e)
            elseif e isa Truss
                append!(bars, ids); push!(bar_mat, mi - 1); push!(areas, area(cross_section(e))); push!(bar_elems, e)
                sc = strain_code(strain_model(e))
                isempty(bar_mat) || length(bar_mat) == 1 || sc == strain ||
                    error("libonsas_cuda evaluates all trusses of a structure with one strain model")
                strain = sc
            else
                error("element type $(typeof(e)) is outside the GPU hot path")
            end
        end
    end
    free0 = Int64.(free_dofs(s)) .- 1
    GC.@preserve xyz kinds params tets tet_mat bars bar_mat areas free0 begin
        h = ctx.handle
        check(ctx, ccall((:onsas_set_nodes, LIB[]), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Float64}), h, n_nodes, n_nodes, dim, xyz))
        check(ctx, ccall((:onsas_set_materials, LIB[]), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Float64}), h, length(kinds), kinds, params))
        isempty(tet_mat) || check(ctx, ccall((:onsas_set_tets, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}), h, length(tet_mat), tets, tet_mat))
        isempty(bar_mat) || check(ctx, ccall((:onsas_set_trusses, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int32),
                                              h, length(bar_mat), bars, bar_mat, areas, strain))
        check(ctx, ccall((:onsas_set_free_dofs, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Int64), h, length(free0), free0, length(free0)))
        check(ctx, ccall((:onsas_finalize_mesh, LIB[]), Int32, (Ptr{Cvoid},), h))
    end
    tet_elems, bar_elems, node_index
end

# mirror of onsas_step_info
struct StepInfo
    norm_dU::Float64; norm_U::Float64; norm_r::Float64; norm_Fext::Float64
    cg_iters::Int64; cg_residual::Float64; cg_tol::Float64; ms_assemble::Float64; ms_solve::Float64
end

"Copy device results into the reference's state so that store!, Solution accessors and write_vtk keep working."
function download!(ctx::CudaContext, state::FullStaticState, tet_elems, bar_elems)
    h = ctx.handle
    U = displacements(state); Fint = internal_forces(state); Fext = external_forces(state)
    GC.@preserve U Fint Fext begin
        check(ctx, ccall((:onsas_get_U, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), h, U))
        check(ctx, ccall((:onsas_get_Fint, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), h, Fint))
        check(ctx, ccall((:onsas_get_Fext, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), h, Fext))   # built on the device when the loads are face loads
    end
    for (family, elems) in ((Int32(0), tet_elems), (Int32(1), bar_elems))
        isempty(elems) && continue
        sig = Matrix{Float64}(undef, 9, length(elems)); eps_ = similar(sig)
        GC.@preserve sig eps_ check(ctx, ccall((:onsas_get_stress_strain, LIB[]), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), h, family, sig, eps_))
        for (k, e) in enumerate(elems)                                         # StructuralAnalyses.jl:115-120
            stress(state)[e] .= reshape(view(sig, :, k), 3, 3)
            strain(state)[e] .= Symmetric(reshape(eps_[:, k], 3, 3))
        end
    end
end

"""
Device-side `apply!` (SURVEY 8f-1): registers one load pattern per `Pressure` BC and one per component of a `GlobalLoad` BC
whose entities are all `TriangularFace`s; returns the closures `t -> factor` in pattern order, or `nothing` when some load
is not a face load (the host `apply!` + `onsas_set_Fext` path is used then).  `node_index` is the 0-BASED node -> id map that
`upload!` returns (the ids the C ABI uses).
"""
function register_loads!(ctx::CudaContext, s::AbstractStructure, node_index::AbstractDict)
    factors = Function[]
    bcs = boundary_conditions(s)
    lbcs = load_bcs(bcs)                             # Vector of load BCs (StructuralBoundaryConditions.jl:188-192); bcs[bc] = its entities
    all(bc -> all(e -> e isa TriangularFace, bcs[bc]), lbcs) || return nothing   # decide before anything is registered
    for bc in lbcs
        ents = bcs[bc]
        tri = Int32[node_index[n] for e in ents for n in nodes(e)]              # 3 x n_faces, already 0-based
        pid = Ref{Int32}(-1)
        if bc isa Pressure
            vals = [1.0, 0.0, 0.0]
            GC.@preserve tri vals check(ctx, ccall((:onsas_add_face_load, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Int32, Ptr{Float64}, Ref{Int32}),
                ctx.handle, length(ents), tri, Int32(1), vals, pid))
            push!(factors, t -> bc(t))
        else
            for c in 1:3
                e_c = [Float64(c == k) for k in 1:3]
                GC.@preserve tri e_c check(ctx, ccall((:onsas_add_face_load, LIB[]), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Int32, Ptr{Float64}, Ref{Int32}),
                    ctx.handle, length(ents), tri, Int32(0), e_c, pid))
                push!(factors, t -> bc(t)[c])
            end
        end
    end
    factors
end

"`apply!(sa, load_bcs)` of this load step: on the device when every load is a face load, else on the host (StructuralAnalyses.jl:228-241)."
function apply_loads!(ctx::CudaContext, sa, factors)
    state = current_state(sa)
    if factors === nothing
        external_forces(state) .= 0
        apply!(sa, load_bcs(boundary_conditions(structure(sa))))
        Fext = external_forces(state)
        GC.@preserve Fext check(ctx, ccall((:onsas_set_Fext, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.handle, Fext))
    else
        f = Float64[g(current_time(sa)) for g in factors]
        GC.@preserve f check(ctx, ccall((:onsas_apply_loads, LIB[]), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}), ctx.handle, length(f), f))
    end
end

"Drop-in replacement of `_solve!(::NonLinearStaticAnalysis, ::AbstractSolver, ...)` (NonLinearStaticAnalyses.jl:70-104)."
function _solve!(sa::NonLinearStaticAnalysis, alg::NewtonRaphsonCUDA, linear_solver::LinearSolver = nothing;
        linear_solve_inplace::Bool = false)
    s = structure(sa)
    ctx = CudaContext(alg.device)                    # created lazily here: `solve` has already deep-copied `sa`
    tet_elems, bar_elems, node_index = upload!(ctx, s)
    factors = register_loads!(ctx, s, node_index)    # device-side F_ext when every load is a face load
    state = current_state(sa)
    U = displacements(state)
    GC.@preserve U check(ctx, ccall((:onsas_set_U, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.handle, U))
    sol = Solution(sa, alg)
    info = Ref{StepInfo}()
    while !is_done(sa)
        step = sa.current_step
        reset!(current_iteration(sa))                                            # :83
        apply_loads!(ctx, sa, factors)                                           # :86-87
        while isconverged!(current_iteration(sa), tolerances(alg)) isa NotConvergedYet   # :90
            # assemble!(s, sa) + step!(sa, alg, linear_solver) in one call (:92-95, :107-148)
            check(ctx, ccall((:onsas_newton_step, LIB[]), Int32, (Ptr{Cvoid}, Int32, Float64, Float64, Int64, Ref{StepInfo}),
                ctx.handle, precond_code(alg), alg.cg_reltol, alg.cg_abstol, alg.cg_maxiter, info))
            i = info[]
            update!(current_iteration(sa), i.norm_dU, i.norm_dU / i.norm_U, i.norm_r, i.norm_r / i.norm_Fext)  # :138-147
        end
        download!(ctx, state, tet_elems, bar_elems)
        store!(sol, state, step)                                                 # :98
        next!(sa)                                                                # :101
    end
    sol
end

"""
Drop-in replacement of `_solve!(::LinearStaticAnalysis, ::Nothing, linear_solver)` (LinearStaticAnalyses.jl:78-113) in the
reference's own sequence: K assembled at the first load step only (:93-96), `step!` (:117-153) = `onsas_step` with
`update_U = 2` (r = F_ext[free], U[free] = dU), then `assemble!` again for stress / strain (:101-102).
Usage: `solve!(LinearStaticAnalysis(s; NSTEPS), NewtonRaphsonCUDA())`.
"""
function _solve!(sa::LinearStaticAnalysis, alg::NewtonRaphsonCUDA, linear_solver::LinearSolver = nothing;
        linear_solve_inplace::Bool = false)
    s = structure(sa)
    ctx = CudaContext(alg.device)
    tet_elems, bar_elems, node_index = upload!(ctx, s)
    factors = register_loads!(ctx, s, node_index)
    state = current_state(sa)
    U0 = copy(displacements(state))
    GC.@preserve U0 check(ctx, ccall((:onsas_set_U, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.handle, U0))
    k_is_constant = isempty(bar_elems) && all(m -> m isa IsotropicLinearElastic, keys(materials(s)))
    sol = Solution(sa, alg)
    info = Ref{StepInfo}()
    while !is_done(sa)
        step = sa.current_step
        apply_loads!(ctx, sa, factors)                                           # :89-90
        if step == 1 || !k_is_constant
            # :93-96 K of the initial state.  The reference solves later steps with the copy it froze at step 1; the device
            # holds one K, so a K that depends on U is re-evaluated at the initial state (same matrix, same result)
            step == 1 || GC.@preserve U0 check(ctx, ccall((:onsas_set_U, LIB[]), Int32, (Ptr{Cvoid}, Ptr{Float64}), ctx.handle, U0))
            check(ctx, ccall((:onsas_assemble, LIB[]), Int32, (Ptr{Cvoid},), ctx.handle))
        end
        check(ctx, ccall((:onsas_step, LIB[]), Int32, (Ptr{Cvoid}, Int32, Float64, Float64, Int64, Int32, Ref{StepInfo}),
            ctx.handle, precond_code(alg), alg.cg_reltol, alg.cg_abstol, alg.cg_maxiter, Int32(2), info))   # :99
        check(ctx, ccall((:onsas_assemble, LIB[]), Int32, (Ptr{Cvoid},), ctx.handle))                       # :101-102
        download!(ctx, state, tet_elems, bar_elems)
        store!(sol, state, step)                                                 # :105
        next!(sa)                                                                # :108
    end
    sol
end

end # module
