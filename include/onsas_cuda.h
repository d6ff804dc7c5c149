/*
 * onsas_cuda.h -- C ABI of libonsas_cuda, the B200-native replacement of ONSAS.jl's
 * Newton-Raphson hot path (per-element f_int / K_t / stress evaluation, global assembly,
 * the linear solve inside each Newton iteration).
 *
 * The reference (ONSAS.jl v0.4.6, pure Julia) has no FFI on this path; it extends by multiple
 * dispatch.  The entry points below are what a `ccall` shim placed at the reference's own seams
 * would bind (INTEGRATION.md shows the Julia side).  file:line citations are relative to the
 * reference's src/ directory.
 *
 * Conventions
 *   - plain pointers and sizes only; all floating point data is FP64, all host pointers;
 *   - every function returns an int32 status (ONSAS_OK = 0); onsas_last_error() gives the text;
 *   - node ids, element ids and dofs are 0-based (the Julia glue subtracts 1 once at upload);
 *   - global dof of component c of node i is dim*i + c  (Meshes/Meshes.jl:85-98 set_dofs!);
 *   - element dofs are node-major (Entities/Entities.jl:156-171 local_dofs);
 *   - 3x3 tensors and K_e are column-major = Julia `Matrix` memory order;
 *   - "stress" is P = F*S and "strain" is C = F'F for hyperelastic tetrahedra
 *     (Entities/Tetrahedrons.jl:218-221), Cauchy stress / small strain for
 *     IsotropicLinearElastic (:265), [1,1]-only for trusses (Entities/Trusses.jl:148-152);
 *   - the caller owns every host buffer; the library copies and never keeps a host pointer;
 *   - a context is used by one host thread at a time; different contexts are independent;
 *   - there is NO CPU fallback: without a CUDA device onsas_create fails with ONSAS_ERR_CUDA.
 */
#ifndef ONSAS_CUDA_H
#define ONSAS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct onsas_ctx onsas_ctx;

/* status codes */
#define ONSAS_OK 0
#define ONSAS_ERR_INVALID_ARG 1
#define ONSAS_ERR_NEGATIVE_VOLUME 2 /* -> ArgumentError("Element with negative volume, check connectivity.") Tetrahedrons.jl:136 */
#define ONSAS_ERR_CUDA 3
#define ONSAS_ERR_NOT_READY 4 /* mesh not finalized / state missing */
#define ONSAS_ERR_UNSUPPORTED 5
#define ONSAS_ERR_COMM 6
#define ONSAS_ERR_ALLOC 7
#define ONSAS_ERR_BREAKDOWN 8 /* the CG residual became non-finite (NaN / Inf in K, F_ext or U): the solve stops at once */

/* material kinds: parameters (p0, p1) */
#define ONSAS_MAT_SVK 0        /* (lambda, G)   Materials/SVKMaterial.jl:25-54 */
#define ONSAS_MAT_NEOHOOKEAN 1 /* (K, G)        Materials/NeoHookeanMaterial.jl:25-56 */
#define ONSAS_MAT_ISOLINEAR 2  /* (E, nu)       Materials/IsotropicLinearElasticMaterial.jl:22-35 */

/* truss strain models (Entities/Trusses.jl:26-45) */
#define ONSAS_STRAIN_ROTATED_ENGINEERING 0
#define ONSAS_STRAIN_GREEN 1

/* element families */
#define ONSAS_FAMILY_TET 0
#define ONSAS_FAMILY_TRUSS 1

/* preconditioners for the CG solve */
#define ONSAS_PRECOND_NONE 0   /* IterativeSolversJL_CG default of the reference, StructuralSolvers.jl:29 */
#define ONSAS_PRECOND_JACOBI 1 /* the north-star solver */
#define ONSAS_PRECOND_TWO_LEVEL 2 /* Jacobi + aggregated coarse space (rigid-body modes of node aggregates: translations, in 3D also
                                     rotations; E = Z^T K Z inverted explicitly once per assembly); streamed persistent solver only.
                                     SURVEY 8f-4 */

/* tuning keys for onsas_set_option */
#define ONSAS_OPT_CG_MODE 1      /* 0 = persistent cooperative kernel with K streamed through shared memory by TMA bulk copies (default;
                                        on N GPUs halo values and partial sums travel over NVLink peer memory),
                                    1 = one launch per phase (NCCL between the launches when N > 1),
                                    2 = persistent cooperative kernel with register-fed SpMV (fallback of 0 for very wide rows) */
#define ONSAS_OPT_ASM_MINBLOCKS 2 /* register budget of the assembly kernel: 1 = unconstrained, 2 = 128, 3 = 96 registers (default 3 = 3 resident CTAs of 192 threads) */
#define ONSAS_OPT_CG_CHECK_EVERY 3 /* multi-launch CG: iterations enqueued between host convergence checks (default 16) */
#define ONSAS_OPT_CG_BLOCKS_PER_SM 4 /* persistent CG: resident CTAs per SM the kernel is compiled for: 4, 5 or 6 (default 6); 1-3 shrink the grid */
#define ONSAS_OPT_CG_PROFILE 5       /* 1 = the persistent CG kernel records per-phase SM-clock cycles (onsas_get_cg_profile) */
#define ONSAS_OPT_FORCE_MG 6         /* diagnostics: set before onsas_finalize_mesh to run the multi-GPU CG kernel with a single rank */
#define ONSAS_OPT_HOST_CHUNKS 7      /* onsas_assemble_host: slice ranges the assembly is cut into so that the copies of U and F_int overlap it (default 12; 1 = no overlap) */
#define ONSAS_OPT_GJ_BLOCKED 8       /* two-level preconditioner: 1 = coarse inverse by 12-row panels (default), 0 = one pivot row per grid barrier */
#define ONSAS_OPT_HOST_MID_WEIGHT 9  /* onsas_assemble_host: size of an inner slice range relative to the first / last one (default 3) */
#define ONSAS_OPT_COARSE_RBM 10      /* two-level preconditioner in 3D: 1 = rigid-body rotations of every aggregate join the coarse space (default), 0 = translations only */
#define ONSAS_OPT_HOST_STREAMS 12     /* onsas_assemble_host: compute streams consecutive slice ranges alternate on (1 or 2, default 2) */
#define ONSAS_OPT_REORDER 13         /* set before onsas_finalize_mesh: 1 = the nodes are renumbered along the Z-curve of their coordinates inside the
                                        library (per device on a multi-device context), so that an arbitrary caller numbering (a Gmsh mesh,
                                        Interfaces/Gmsh.jl:25-83) gets the slice locality of a structured one; invisible to the caller: every vector,
                                        dof list, face list and result stays in the caller's numbering.  2 = aggregate-major: the nodes of every aggregate of the
                                        two-level preconditioner are numbered consecutively (Z-curve inside an aggregate), so its aggregate-ordered
                                        passes stream through memory.  0 (default) keeps the caller's order */
#define ONSAS_OPT_CG_SINGLE_REDUCTION 14 /* streamed persistent solver, bit mask: bit 0 (default on) = Jacobi-PCG runs the single-reduction recurrence
                                            (Chronopoulos-Gear: one reduction of (r.r, r.u, u.Ku) and two grid barriers per iteration, one cross-GPU
                                            all-reduce; the classic recurrence has three barriers and two all-reduces); bit 1 (default off) = the
                                            two-level PCG too (three barriers instead of four, but more vector traffic: measured slower per iteration,
                                            profiles/r48/sr_ab.log).  Same iterates in exact arithmetic; precond = 0 always runs the classic
                                            recurrence (IterativeSolvers' cg! step by step) */
#define ONSAS_OPT_TRUSS_MINBLOCKS 15  /* register budget of the truss assembly kernel: 2 = 120, 3 = 80, 4 = 64 registers (default: 8.2 G bars/s on the 10 M-bar lattice against 6.5 at 120) */
#define ONSAS_OPT_HOST_GRAPH 16      /* onsas_assemble_host: 1 = when the same pinned buffers are passed twice in a row, the pipeline of copies and kernels is
                                        captured once into a CUDA graph and every later call is one graph launch; 0 (default) = always enqueue it
                                        (measured on B200: the replayed 80-node, 4-stream graph takes 0.409 ms per pass, the eager enqueue 0.383) */
#define ONSAS_OPT_COARSE_GLOBAL 17   /* two-level preconditioner on several GPUs (3-D, native partition, peer-memory solver): 1 = a global coarse level
                                        couples the ranks (level-2 aggregates = unions of each rank's aggregates, E2 assembled by all ranks over the peer
                                        window, folded into each rank's dense operator: one more L2-resident dense apply and one all-gather of <= 1536
                                        doubles per iteration); 0 (default) = every rank's coarse space stands alone (block-diagonal E).  Measured
                                        (profiles/r61, r62): correct, but the extra latency per iteration outweighs the iterations it saves */
#define ONSAS_OPT_CG_L2_PREFETCH 18  /* streamed persistent solver: slices per consumer warp the producer warp pulls into L2 (cp.async.bulk.prefetch.L2)
                                        behind the shared-memory ring as soon as an SpMV phase has issued its last copy -- HBM idles during the vector
                                        phases and grid barriers of an iteration, so the first part of the next SpMV phase is served from L2.
                                        0 = off (default).  Measured on B200 (profiles/r64, r66): bitwise identical and slower at every distance --
                                        the prefetched lines displace the CG vectors from L2; kept as the record of the experiment */
#define ONSAS_OPT_COARSE_FUSED 11    /* two-level preconditioner: 1 = residual update in aggregate order, fused with w = Z^T r (default), 0 = separate pass */

/* ---------------------------------------------------------------- life cycle */

/* Creates a context on CUDA device `device` (replaces nothing in the reference: the device
 * state is the analogue of FullStaticState, StructuralAnalyses/StaticStates.jl:33-102). */
int32_t onsas_create(int32_t device, onsas_ctx** ctx);
int32_t onsas_destroy(onsas_ctx* ctx);
const char* onsas_last_error(onsas_ctx* ctx); /* ctx may be NULL: last error of onsas_create */
int32_t onsas_version(void);
/* Run all work on the given cudaStream_t (e.g. torch's current stream); NULL = the context's own stream. */
int32_t onsas_set_stream(onsas_ctx* ctx, void* cuda_stream);
int32_t onsas_set_option(onsas_ctx* ctx, int32_t key, int64_t value);

/* ---------------------------------------------------------------- mesh upload (once per Structure) */

/* xyz: dim x n_nodes column-major (node-interleaved).  Entities/Nodes.jl:74-84, Meshes/Meshes.jl:158-184.
 * Multi-GPU: the first n_owned nodes are the block rows this rank owns, the rest are halo nodes
 * (columns only); single GPU: n_owned = n_nodes. */
int32_t onsas_set_nodes(onsas_ctx* ctx, int64_t n_nodes, int64_t n_owned, int32_t dim, const double* xyz);
/* kind[n], params 2 x n column-major.  StructuralModel/StructuralMaterials.jl:25-47. */
int32_t onsas_set_materials(onsas_ctx* ctx, int32_t n, const int32_t* kind, const double* params);
/* conn: 4 x n column-major node ids; mat_id[n] or NULL (= material 0).  Entities/Tetrahedrons.jl:24-36. */
int32_t onsas_set_tets(onsas_ctx* ctx, int64_t n, const int32_t* conn, const int32_t* mat_id);
/* conn: 2 x n; area[n] = area(cross_section(e)).  Entities/Trusses.jl:54-72. */
int32_t onsas_set_trusses(onsas_ctx* ctx, int64_t n, const int32_t* conn, const int32_t* mat_id, const double* area,
                          int32_t strain_model);
/* free dofs of the owned nodes (Structures.jl:129-142 free_dofs), any order, local numbering (multi-GPU: free dofs of
 * halo nodes may be listed too, so that the solver updates U on them exactly like their owner does).
 * n_free_global = number of free dofs of the whole structure (= n_free on one GPU); it is the CG
 * default maxiter (StructuralSolvers.jl:229-234). */
int32_t onsas_set_free_dofs(onsas_ctx* ctx, int64_t n_free, const int64_t* free_dofs, int64_t n_free_global);
/* Builds the sparsity pattern, the row-owner pair lists and contribution lists, uploads everything,
 * checks element volumes (Tetrahedrons.jl:134-138).  Replaces Assembler construction and the
 * first end_assemble! (Assemblers.jl:28-40, 84-88; StaticStates.jl:79-102). */
int32_t onsas_finalize_mesh(onsas_ctx* ctx);

/* ---------------------------------------------------------------- state vectors (n_local_dofs = dim*n_nodes) */

int32_t onsas_set_U(onsas_ctx* ctx, const double* U);       /* displacements(state), StructuralAnalyses.jl:74 */
int32_t onsas_get_U(onsas_ctx* ctx, double* U);
int32_t onsas_set_Fext(onsas_ctx* ctx, const double* Fext); /* external_forces(state), :88; result of apply! :228-241 */
int32_t onsas_get_Fint(onsas_ctx* ctx, double* Fint);       /* internal_forces(state), :85; reactions = Fint at fixed dofs */
int32_t onsas_get_Fext(onsas_ctx* ctx, double* Fext);
int32_t onsas_get_dU(onsas_ctx* ctx, double* dU);           /* Delta_displacements(state) scattered to a full dof vector */

/* ---------------------------------------------------------------- external loads on the device
 * The step before every Newton loop (apply!(sa, load_bcs), StructuralAnalyses.jl:228-241 with
 * StructuralBoundaryConditions.jl:195-220): each load boundary condition becomes a unit nodal vector ("pattern")
 * built ONCE on the device from the boundary faces; per load step only the factors p_k(t) cross the boundary and
 * F_ext = sum_k factors[k] * pattern_k is formed on the device (fixed order: deterministic).
 *   onsas_add_face_load  kind 0 = GlobalLoad on TriangularFaces (GlobalLoadBoundaryConditions.jl:50-68):
 *                        values[3] * A/3 per face node; kind 1 = Pressure (LocalLoadBoundaryConditions.jl:36-56):
 *                        -n * A/3 * values[0], n A = 1/2 (x2-x1) x (x3-x1) (TriangularFaces.jl:44-65).
 *                        tri = 3 x n_faces local 0-based node ids.  A time-dependent direction is expressed as one
 *                        pattern per component (values = e_x, e_y, e_z) with factors = the components of p(t).
 *   onsas_add_nodal_load GlobalLoad on nodes (:34-48): values[dim] on every listed node, duplicates summed.
 *   onsas_apply_loads    n_factors must equal the number of patterns added since the last finalize / clear. */
int32_t onsas_add_face_load(onsas_ctx* ctx, int64_t n_faces, const int32_t* tri, int32_t kind, const double* values,
                            int32_t* pattern_id);
int32_t onsas_add_nodal_load(onsas_ctx* ctx, int64_t n, const int32_t* nodes, const double* values, int32_t* pattern_id);
int32_t onsas_apply_loads(onsas_ctx* ctx, int32_t n_factors, const double* factors);
int32_t onsas_clear_loads(onsas_ctx* ctx);

/* ---------------------------------------------------------------- hot path */

/* assemble!(s, sa): reset, evaluate every element at the current U, assemble F_int, K, stress, strain.
 * StructuralAnalyses/StaticAnalyses.jl:99-132.  Multi-GPU: refreshes the halo part of U first. */
int32_t onsas_assemble(onsas_ctx* ctx);

/* assemble!(s, sa) with the state on the HOST on both sides, the way the reference's caller holds it
 * (`displacements(state)` in, `internal_forces(state)` out; StaticAnalyses.jl:99-122, StaticStates.jl:33-102):
 * U (dim * n_local_nodes doubles) is copied in, K / F_int / stress / strain are assembled on the device and F_int
 * (same length; zero on halo nodes) is copied out.  Equivalent to onsas_set_U + onsas_assemble + onsas_get_Fint and bitwise
 * identical in every result, but the two copies are pipelined with the kernel over ONSAS_OPT_HOST_CHUNKS slice ranges
 * (pinned host buffers are needed for the overlap, not for correctness).  Synchronous: both buffers are free on return. */
int32_t onsas_assemble_host(onsas_ctx* ctx, const double* U, double* F_int);

/* internal_forces(mat, e, u_e[, cache]) for elements [first, first+count) of one family, un-assembled:
 * f (ndof_e per element), K (ndof_e^2, column-major), sig, eps (9 each, column-major).
 * Entities/Entities.jl:174-218; Tetrahedrons.jl:186-266; Trusses.jl:126-184. */
int32_t onsas_eval_elements(onsas_ctx* ctx, int32_t family, int64_t first, int64_t count, double* f, double* K,
                            double* sig, double* eps);

typedef struct onsas_step_info {
    double norm_dU;   /* ||dU||                    NonLinearStaticAnalyses.jl:138 */
    double norm_U;    /* ||U|| BEFORE the update   :139 */
    double norm_r;    /* ||r|| at the START of the iteration :140 */
    double norm_Fext; /* ||F_ext|| (full vector)   :141 */
    int64_t cg_iters;
    double cg_residual; /* final ||r_cg|| */
    double cg_tol;      /* max(reltol*||r0||, abstol) */
    double ms_assemble; /* device time of the assembly of this call (CUDA events) */
    double ms_solve;    /* device time of residual + PCG + update */
} onsas_step_info;

/* One Newton iteration = assemble! + step! (NonLinearStaticAnalyses.jl:92-95, 107-148):
 * assemble at U, r = (F_ext - F_int)[free], solve K[free,free] dU = r by (P)CG with
 * tolerance max(cg_reltol*||r||, cg_abstol) and at most cg_maxiter iterations (<= 0: n_free, the
 * reference default StructuralSolvers.jl:229-234), U[free] += dU, norms. */
int32_t onsas_newton_step(onsas_ctx* ctx, int32_t precond, double cg_reltol, double cg_abstol, int64_t cg_maxiter,
                          onsas_step_info* info);

/* step! without the assembly: residual, solve, update, norms on the K / F_int already assembled.
 * update_U = 0 leaves U untouched (dU still available through onsas_get_dU); 1: U[free] += dU (the Newton update);
 * 2: the step! of a LinearStaticAnalysis (LinearStaticAnalyses.jl:117-153): r = F_ext[free] (F_int is not subtracted),
 * U[free] = dU. */
int32_t onsas_step(onsas_ctx* ctx, int32_t precond, double cg_reltol, double cg_abstol, int64_t cg_maxiter,
                   int32_t update_U, onsas_step_info* info);

/* The linear solve alone: K[free,free] x = b[free] on the assembled K, zero initial guess
 * (solve(LinearProblem(A,b), IterativeSolversJL_CG(); abstol, reltol, maxiter), NonLinearStaticAnalyses.jl:129-134).
 * b, x: full dof vectors (entries at fixed dofs ignored / zero). */
int32_t onsas_pcg(onsas_ctx* ctx, const double* b, double* x, int32_t precond, double reltol, double abstol,
                  int64_t maxiter, int64_t* iters, double* residual);

/* y = K[free,free] * x with full-length vectors (zeros at fixed dofs): one SpMV launch, exposed for tests/bench. */
int32_t onsas_spmv(onsas_ctx* ctx, const double* x, double* y);

/* Measurement hooks (asynchronous, device-resident data only): one SpMV launch on whatever the search
 * direction buffer currently holds; and a stream synchronisation that also reports deferred errors
 * (negative element volume seen by an asynchronous onsas_assemble). */
int32_t onsas_spmv_resident(onsas_ctx* ctx);
int32_t onsas_synchronize(onsas_ctx* ctx);

/* ---------------------------------------------------------------- results */

/* tangent_matrix(state) as scalar CSR with sorted columns: n_rows = dim*n_owned rows over dim*n_nodes columns,
 * fixed dofs included (StaticStates.jl:84-88). */
int32_t onsas_get_csr_size(onsas_ctx* ctx, int64_t* n_rows, int64_t* nnz);
int32_t onsas_get_csr(onsas_ctx* ctx, int64_t* rowptr, int32_t* col, double* val);
/* stress(state)[e], strain(state)[e] for a whole family: 9 doubles per element each, column-major.
 * StructuralAnalyses.jl:115-120; StaticAnalyses.jl:157-174 (store!). */
int32_t onsas_get_stress_strain(onsas_ctx* ctx, int32_t family, double* sig, double* eps);

/* Sizes of the device-side tables, for roofline accounting: out[0]=n_slices, [1]=padded block slots,
 * [2]=true nonzero blocks, [3]=tet pairs, [4]=truss pairs, [5]=max pairs per slice, [6]=bytes of K values,
 * [7]=persistent CG grid size. */
int32_t onsas_get_table_stats(onsas_ctx* ctx, int64_t out[8]);

/* Diagnostics: cycles block 0 of the last persistent CG solve spent in [0] p update, [1] grid sync, [2] SpMV + dot,
 * [3] grid sync, [4] x/r update + dots, [5] grid sync, [6] scalar reductions (needs ONSAS_OPT_CG_PROFILE = 1). */
int32_t onsas_get_cg_profile(onsas_ctx* ctx, int64_t out[8]);

/* ---------------------------------------------------------------- multi-GPU, one process driving all devices
 *
 * The reference's solve is a single process (StructuralSolvers.jl:201-222); a multi-device context keeps that calling
 * convention on N GPUs of one box.  onsas_create_multi returns an ordinary onsas_ctx*: EVERY entry point above takes the
 * GLOBAL mesh and GLOBAL dof vectors in the caller's own numbering, as on one device.  onsas_finalize_mesh partitions the
 * mesh element-wise (recursive coordinate bisection of the nodes into contiguous owned ranges, interface elements
 * evaluated on both sides, halo nodes as extra columns), uploads one part per device, enables peer access and wires the
 * peer-memory solver; one host thread then drives all devices (every call enqueues all devices before it waits for any).
 * Not available on such a context: onsas_set_stream, ONSAS_OPT_CG_MODE = 1, the one-process-per-GPU calls below. */
int32_t onsas_create_multi(const int32_t* devices, int32_t ndev, onsas_ctx** ctx);
int32_t onsas_device_count(onsas_ctx* ctx); /* 1 for onsas_create, ndev for onsas_create_multi */

/* The partitioner on its own, for the one-process-per-GPU binding: every process builds the same partition of the global
 * mesh (deterministic) and loads ITS rank into its context -- onsas_part_load = onsas_set_nodes + onsas_set_tets +
 * onsas_set_trusses + onsas_set_free_dofs + onsas_set_halo of that rank (materials are set separately, before
 * onsas_finalize_mesh).  onsas_part_local_to_global gives, per local node (owned first, then halo), the caller's node id:
 * local vectors are v_local[dim*i + c] = v_global[dim*l2g[i] + c].  onsas_part_sizes: out = {n_local_nodes, n_owned_nodes,
 * n_local_tets, n_local_trusses, n_free_local, n_neighbours, n_send_nodes, n_free_global}.  A context loaded from a
 * partition knows the remote halo offsets itself: pass NULL for them to onsas_p2p_import. */
typedef struct onsas_part onsas_part;
int32_t onsas_part_create(int32_t dim, int64_t n_nodes, const double* xyz, int64_t n_tets, const int32_t* tets, const int32_t* tet_mat,
                          int64_t n_trusses, const int32_t* trusses, const int32_t* truss_mat, const double* area, int64_t n_free,
                          const int64_t* free_dofs, int32_t n_ranks, int32_t reorder /* see ONSAS_OPT_REORDER */, onsas_part** part);
int32_t onsas_part_destroy(onsas_part* part);
int32_t onsas_part_sizes(onsas_part* part, int32_t rank, int64_t out[8]);
int32_t onsas_part_local_to_global(onsas_part* part, int32_t rank, int32_t* l2g);
int32_t onsas_part_local_elements(onsas_part* part, int32_t rank, int32_t family, int64_t* global_ids);
int32_t onsas_part_halo_plan(onsas_part* part, int32_t rank, int32_t* nbr_rank, int64_t* send_ptr, int32_t* send_nodes,
                             int64_t* recv_ptr, int64_t* remote_halo_off); /* any pointer may be NULL */
int32_t onsas_part_load(onsas_part* part, int32_t rank, onsas_ctx* ctx, int32_t truss_strain_model);

/* ---------------------------------------------------------------- multi-GPU (one process per GPU) */

/* NCCL bootstrap: rank 0 calls onsas_comm_unique_id and the host broadcasts the 128 bytes
 * (torch.distributed / MPI); every rank then calls onsas_comm_init. */
int32_t onsas_comm_unique_id(void* id128);
int32_t onsas_comm_init(onsas_ctx* ctx, int32_t n_ranks, int32_t rank, const void* id128);
/* Halo plan: for neighbour k (rank nbr_rank[k]) send the values of owned nodes
 * send_nodes[send_ptr[k] .. send_ptr[k+1]) and receive into the halo nodes
 * [n_owned + recv_ptr[k], n_owned + recv_ptr[k+1]) (halo nodes are grouped by owner). */
int32_t onsas_set_halo(onsas_ctx* ctx, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_ptr,
                       const int32_t* send_nodes, const int64_t* recv_ptr);

/* Peer-memory path of the multi-GPU linear solve: one persistent kernel per GPU; the z values of interface dofs and
 * the scalar partial sums are stored straight into the other ranks' memory over NVLink (8-byte words carrying
 * 32 data bits + a 32-bit epoch, polled by the receiver), no NCCL call inside the solve.
 * After onsas_comm_init + onsas_set_halo + onsas_finalize_mesh every rank exports its window (a 64-byte CUDA IPC
 * handle + the offset of the window inside that allocation); the host all-gathers them and calls onsas_p2p_import
 * with all handles / offsets (indexed by rank) and, per neighbour k, the position (in nodes) inside THAT
 * neighbour's halo where this rank's values start (its receive offset recv_ptr[j] for this rank).
 * Without the import the solve uses the NCCL multi-launch path. */
int32_t onsas_p2p_export(onsas_ctx* ctx, void* handle64, int64_t* offset);
int32_t onsas_p2p_import(onsas_ctx* ctx, const void* handles, const int64_t* offsets, const int64_t* remote_halo_off);

#ifdef __cplusplus
}
#endif
#endif /* ONSAS_CUDA_H */
