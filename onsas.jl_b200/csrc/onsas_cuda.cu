// onsas_cuda.cu -- context, host orchestration and the C ABI of libonsas_cuda (include/onsas_cuda.h).
// There is no CPU execution path in this library: every entry point that computes launches CUDA
// kernels on the context's device and fails with ONSAS_ERR_CUDA when that is impossible.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/onsas_cuda.h"
#include "kernels.cuh"
#include "partition.hpp"
#include "tables.hpp"

using namespace onsas;

namespace {

struct OnsasError : std::runtime_error {
    int code;
    OnsasError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            throw OnsasError(ONSAS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +   \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + ")");         \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;
    void alias(T* ptr, size_t count) {  // view into memory owned elsewhere (the P2P window)
        release();
        p = ptr;
        n = count;
        owned = false;
    }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) {
            cudaError_t e = cudaMalloc(&p, count * sizeof(T));
            if (e != cudaSuccess) {
                p = nullptr;
                n = 0;
                throw OnsasError(ONSAS_ERR_ALLOC, std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) +
                                                      " bytes failed: " + cudaGetErrorString(e));
            }
        }
    }
    void zero(cudaStream_t s) {
        if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void upload(const T* h, size_t count, cudaStream_t s) {
        if (count != n) alloc(count);
        if (count) CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    ~DevBuf() { release(); }
};

// ---- NCCL through dlopen so that single-GPU use has no NCCL dependency
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            err = std::string("cannot load libnccl: ") + dlerror();
            return false;
        }
#define LOAD(sym)                                                    \
    *(void**)(&sym) = dlsym(h, "nccl" #sym);                         \
    if (!sym) {                                                      \
        err = "libnccl lacks nccl" #sym;                             \
        return false;                                                \
    }
        LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(Send) LOAD(Recv) LOAD(GroupStart)
        LOAD(GroupEnd) LOAD(GetErrorString)
#undef LOAD
        return true;
    }
};
NcclApi g_nccl;
std::string g_create_error;

#define NCCL_CHECK(expr)                                                                                    \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess)                                                                              \
            throw OnsasError(ONSAS_ERR_COMM, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));        \
    } while (0)

}  // namespace

struct onsas_ctx {
    int device = 0;
    int n_sm = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;

    // host copy of the mesh until finalize
    int dim = 3;
    int64_t n_nodes = 0, n_owned = 0;
    std::vector<double> h_xyz;
    std::vector<int32_t> h_tets, h_tet_mat, h_trusses, h_truss_mat;
    bool tet_has_mat = false, truss_has_mat = false;
    std::vector<double> h_area;
    int strain_model = 0;
    std::vector<int32_t> h_mat_kind;
    std::vector<double> h_mat_params;
    std::vector<uint8_t> h_mask;
    int64_t n_free = 0, n_free_global = 0;
    bool have_nodes = false, have_free = false, finalized = false;
    int64_t n_tets = 0, n_trusses = 0;
    int tet_kind = MAT_SVK;  // uniform kind or MAT_MIXED

    MeshTables tab;

    // device
    DevBuf<double> X, U, Fext, Fint, val, x, r, p, p_pad, Ap, dinv, rhs, s_vec, partials, red, tet_out, truss_out, area, mat_params;
    DevBuf<int32_t> tets, tet_mat, trusses, truss_mat, mat_kind, col, diag_slot, pair_code[2], snodes[2], send_nodes;
    DevBuf<uint16_t> pair_lnodes[2];
    DevBuf<int64_t> slice_ptr;
    DevBuf<SliceHdr> hdr[2];
    DevBuf<uint32_t> cptr[2];
    DevBuf<uint16_t> ccode[2];
    DevBuf<uint8_t> mask;
    DevBuf<int> err_flag;
    DevBuf<CgState> st;
    DevBuf<double> sendbuf;
    DevBuf<long long> prof;
    int cg_profile = 0;
    CgState* h_st = nullptr;  // pinned
    int* h_flag = nullptr;    // pinned
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};

    // options
    int cg_mode = 0, asm_minb = 3, truss_minb = 4, check_every = 16, cg_bps = 6;
    int cg_grid = 0, part_stride = 4096;
    // two-level preconditioner (precond = 2): node aggregates, dense coarse inverse, work vectors
    struct Coarse {
        bool built = false;   // aggregates exist for the current mesh
        bool fresh = false;   // Einv matches the K currently assembled
        int n_agg = 0, nc = 0, cd = 0;
        bool glob = false;    // the global level is part of the operator
        int nc2 = 0, n2_own = 0, agg2_first = 0;
    } co;
    bool coarse_fused = true;  // residual update in aggregate order, fused with w = Z^T r (same-box sweep: profiles/r18)
    bool coarse_rbm = true;  // 3D: rigid-body rotations of every aggregate join the coarse space (6 coarse dofs per aggregate)
    DevBuf<int32_t> co_agg, co_agg_ptr, co_agg_nodes;
    std::vector<double> h_elem_stage;  // onsas_get_stress_strain: host copy of the tets' 16-double records
    DevBuf<double> co_E, co_w, co_y, co_rowbuf, co_rho;
    // global coarse level across ranks (ONSAS_OPT_COARSE_GLOBAL): level-2 aggregates from the partitioner
    std::vector<int32_t> h_agg2;   // [n_local] global level-2 aggregate of every local node (owned + halo)
    std::vector<double> h_cen2;    // [n_agg2_total * dim]
    int n_agg2_per_rank = 0;
    int coarse_global = 0;         // ONSAS_OPT_COARSE_GLOBAL (off by default: measured, per iteration it costs more than its iterations save)
    unsigned long long co2_epoch = 0;
    DevBuf<int32_t> co2_agg, co2_ptr, co2_nodes, co_parent, co_child_ptr;
    DevBuf<double> co2_rho, co_dvec, co_G, co_w2plain, co_rowbuf2;
    DevBuf<double*> d_peer_E2;
    DevBuf<unsigned long long*> d_peer_rowflags, d_peer_w2;
    // onsas_assemble_host: slice ranges launched one after the other while the copies of U (in) and F_int (out) overlap them
    int64_t asm_first = 0, asm_count = -1;  // slice range of the next assembly launches (-1: all slices)
    cudaStream_t asm_stream = nullptr;      // stream of the next assembly launches (null: the context's stream)
    int host_streams = 2;                   // onsas_assemble_host: compute streams the slice ranges alternate on
    int host_chunks = 12, host_mid_weight = 3;  // with two compute streams (sweep: profiles/r29, r30)
    bool gj_blocked = true;  // coarse inverse by the panel (blocked) Gauss-Jordan kernel; false: one pivot row per grid barrier
    struct HostPlan {
        bool built = false;
        std::vector<int64_t> slice0;   // [n_chunks + 1]
        std::vector<int64_t> node_hi;  // [n_chunks] the chunk's elements touch nodes [0, node_hi) only
        cudaStream_t s_in = nullptr, s_out = nullptr, s_k2 = nullptr;
        cudaEvent_t ev_start = nullptr;
        std::vector<cudaEvent_t> ev_in, ev_k;
        cudaEvent_t ev_join[2] = {nullptr, nullptr};
        cudaGraphExec_t gexec = nullptr;        // the captured pipeline for the buffers (gU, gF)
        const double* gU = nullptr;
        double* gF = nullptr;
        const double* seenU = nullptr;          // buffers of the previous call (a pair seen twice in a row is captured)
        double* seenF = nullptr;
        bool graph_failed = false;
    } hp;
    int host_graph = 0;  // ONSAS_OPT_HOST_GRAPH (measured: the replayed graph is slower than the eager enqueue, 0.409 vs 0.383 ms)
    struct StreamPlan {
        bool built = false, ok = false;
        int n_cw = 0, depth = 0, grid = 0, threads = 0;
        void* kern = nullptr;     // classic recurrence (every preconditioner)
        void* kern_sr = nullptr;  // single-reduction recurrence (precond 0 / 1)
        size_t smem = 0;
    } st_plan;
    int cg_single_reduction = 1;  // ONSAS_OPT_CG_SINGLE_REDUCTION: bit 0 = Jacobi-PCG (default), bit 1 = two-level PCG run the single-reduction recurrence
    int force_mg = 0;  // diagnostics: run the multi-GPU kernel even with one rank
    int cg_l2_prefetch = 0;  // ONSAS_OPT_CG_L2_PREFETCH: slices per consumer warp pulled into L2 behind the ring between SpMV phases
    int reorder = 0;   // ONSAS_OPT_REORDER: 1 = the nodes are renumbered along a Z-curve inside onsas_finalize_mesh (invisible to the caller)
    std::vector<std::pair<int32_t, int64_t>> opt_log;  // options in the order they were set (replayed on the device contexts of a group)

    // comm
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    std::vector<int32_t> nbr_rank;
    std::vector<int64_t> send_ptr, recv_ptr;
    // peer-memory window (multi-GPU persistent CG): [slots][flags][epochs][p]
    DevBuf<unsigned char> window;
    std::vector<void*> ipc_opened;
    bool p2p_ready = false;
    std::vector<int32_t> h_send_nodes;
    std::vector<int32_t> h_agg_ptr;        // aggregate-major numbering: owned-node ranges of the preconditioner's aggregates (else empty)
    std::vector<int64_t> remote_halo_off;  // per neighbour: where this rank's values start inside ITS halo (contexts loaded from a partition)
    // device-side load patterns (unit nodal vectors of the load boundary conditions), n_local_dofs each
    DevBuf<double> patterns, factors;
    int n_patterns = 0;
    std::vector<uint8_t> h_iface;  // per owned dof: 1 = a neighbour rank needs its value (mask bit 1 on the device)
    DevBuf<long long> d_push_ptr;
    DevBuf<unsigned long long*> d_push_dst, d_peer_slots;


    // one process driving several devices (onsas_create_multi): this context holds the GLOBAL mesh and vectors in the
    // caller's numbering and owns one ordinary context per device (group.inc)
    struct Group* grp = nullptr;

    int64_t n_local_dofs() const { return n_nodes * dim; }
    int64_t n_own_dofs() const { return n_owned * dim; }
};

// pinned host staging buffer (grown on demand): the copies of a group's local vectors overlap the kernels like the caller's own
struct HostBuf {
    double* p = nullptr;
    size_t n = 0, cap = 0;
    void resize(size_t count) {
        if (count > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            if (cudaMallocHost(&p, count * sizeof(double)) != cudaSuccess) {
                p = nullptr;
                cap = n = 0;
                throw std::bad_alloc();
            }
            cap = count;
        }
        n = count;
    }
    double* data() { return p; }
    size_t size() const { return n; }
    double& operator[](size_t i) { return p[i]; }
    HostBuf() = default;
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
    HostBuf(HostBuf&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
    ~HostBuf() {
        if (p) cudaFreeHost(p);
    }
};

struct Group {
    std::vector<onsas_ctx*> sub;          // one context per device, rank r = sub[r]
    Partition part;                       // the global partition (freed down to what the vector traffic needs after finalize)
    std::vector<LocalPart> lp;            // per rank: l2g, element ids, sizes (connectivity arrays released after upload)
    std::vector<HostBuf> hA, hB;          // per-rank PINNED host staging of local vectors
    int64_t n_free = 0;
};

namespace {

template <typename F>
int32_t guard(onsas_ctx* ctx, F&& f) {
    try {
        if (ctx) {
            cudaError_t e = cudaSetDevice(ctx->device);
            if (e != cudaSuccess) throw OnsasError(ONSAS_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        }
        f();
        return ONSAS_OK;
    } catch (const OnsasError& e) {
        if (ctx) ctx->err = e.what();
        else g_create_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "host allocation failed";
        return ONSAS_ERR_ALLOC;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = e.what();
        return ONSAS_ERR_INVALID_ARG;
    } catch (...) {
        if (ctx) ctx->err = "unknown error";
        return ONSAS_ERR_INVALID_ARG;
    }
}

void require(bool cond, int code, const char* msg) {
    if (!cond) throw OnsasError(code, msg);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (kernel, DEVICE) and is shared by every context of the
// process: keep a running maximum per (device, kernel) under a lock, so that a second context -- on the same device with
// a narrower mesh, or on another device -- can neither lower the limit under a live context nor skip setting it.
void ensure_dyn_smem(int device, const void* kern, size_t smem) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> configured;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = configured[std::make_pair(device, kern)];
    if (smem > have) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
}

// device mask = bit 0: free dof (onsas_set_free_dofs), bit 1: interface dof (onsas_p2p_import)
void upload_mask(onsas_ctx* c) {
    std::vector<uint8_t> m(c->h_mask);
    for (size_t i = 0; i < c->h_iface.size() && i < m.size(); ++i)
        if (c->h_iface[i]) m[i] |= 2;
    c->mask.upload(m, c->stream);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// material ids in range, the element kind(s) in use (one kernel instantiation per uniform kind), truss materials hyperelastic
void derive_element_kinds(onsas_ctx* c) {
    const int nm = (int)c->h_mat_kind.size();
    auto check_mat = [&](const std::vector<int32_t>& ids, int64_t n) {
        for (int64_t e = 0; e < (int64_t)ids.size() && e < n; ++e)
            require(ids[e] >= 0 && ids[e] < nm, ONSAS_ERR_INVALID_ARG, "element material id out of range");
    };
    if (c->tet_has_mat) check_mat(c->h_tet_mat, c->n_tets);
    if (c->truss_has_mat) check_mat(c->h_truss_mat, c->n_trusses);
    // kinds in use
    if (c->n_tets > 0) {
        int kind = -1;
        bool mixed = false;
        if (!c->tet_has_mat) kind = c->h_mat_kind[0];
        else
            for (int64_t e = 0; e < c->n_tets; ++e) {
                int k = c->h_mat_kind[c->h_tet_mat[e]];
                if (kind < 0) kind = k;
                else if (k != kind) { mixed = true; break; }
            }
        c->tet_kind = mixed ? MAT_MIXED : kind;
    }
    if (c->n_trusses > 0) {
        for (int64_t e = 0; e < c->n_trusses; ++e) {
            int k = c->h_mat_kind[c->truss_has_mat ? c->h_truss_mat[e] : 0];
            // Trusses.jl:126,159 dispatch on AbstractHyperElasticMaterial only
            require(k != MAT_ISOLINEAR, ONSAS_ERR_UNSUPPORTED, "trusses need a hyperelastic material (SVK / NeoHookean)");
        }
    }
}

// ---------------------------------------------------------------- assembly launch
template <int FAMILY, int KIND, int DIM, bool ACCUM, int MAXT, int MINB>
void launch_asm_inst(onsas_ctx* c, const AsmArgs& A, int threads, size_t smem) {
    auto kern = k_assemble<FAMILY, KIND, DIM, ACCUM, MAXT, MINB>;
    ensure_dyn_smem(c->device, (const void*)kern, smem);
    kern<<<(unsigned)(c->asm_count < 0 ? c->tab.n_slices : c->asm_count), threads, smem, c->asm_stream ? c->asm_stream : c->stream>>>(A);
    CUDA_CHECK(cudaGetLastError());
}

template <int FAMILY, int KIND, int DIM, bool ACCUM, int REGS>
void launch_asm_reg(onsas_ctx* c, const AsmArgs& A, int threads, size_t smem) {
    auto kern = k_assemble_reg<FAMILY, KIND, DIM, ACCUM, REGS>;
    ensure_dyn_smem(c->device, (const void*)kern, smem);
    kern<<<(unsigned)(c->asm_count < 0 ? c->tab.n_slices : c->asm_count), threads, smem, c->asm_stream ? c->asm_stream : c->stream>>>(A);
    CUDA_CHECK(cudaGetLastError());
}

template <int KIND>
void launch_asm_tets_kind(onsas_ctx* c, const AsmArgs& A, int threads, size_t smem) {
    // register budget variants: 1 -> unconstrained, 2 -> 128 regs (launch bounds 256 x 2), 3 -> 96 regs (3 CTAs of 192: 18 warps need <= 102 regs each, the register file is split per SM sub-partition)
    if (c->asm_minb == 1) launch_asm_inst<0, KIND, 3, false, 256, 1>(c, A, threads, smem);
    else if (c->asm_minb == 3 && threads <= 192) launch_asm_reg<0, KIND, 3, false, 96>(c, A, threads, smem);
    else launch_asm_inst<0, KIND, 3, false, 256, 2>(c, A, threads, smem);
}

AsmArgs make_asm_args(onsas_ctx* c, int family) {
    AsmArgs A{};
    A.n_rows_guard = c->n_owned;
    A.X = c->X.p;
    A.U = c->U.p;
    A.hdr = c->hdr[family].p;
    A.snodes = c->snodes[family].p;
    A.pair_lnodes = c->pair_lnodes[family].p;
    A.max_pairs = std::max(c->tab.fam[family].max_pairs_per_slice, 1);
    A.max_width = std::max(c->tab.max_width, 1);
    A.max_snodes = std::max(c->tab.fam[family].max_snodes, 1);
    A.mat_id = family == 0 ? (c->tet_has_mat ? c->tet_mat.p : nullptr) : (c->truss_has_mat ? c->truss_mat.p : nullptr);
    A.mat_kind = c->mat_kind.p;
    A.mat_params = c->mat_params.p;
    A.area = c->area.p;
    A.strain_model = c->strain_model;
    A.pair_code = c->pair_code[family].p;
    A.cptr = c->cptr[family].p;
    A.ccode = c->ccode[family].p;
    A.val = c->val.p;
    A.F_int = c->Fint.p;
    A.elem_out = family == 0 ? c->tet_out.p : c->truss_out.p;
    A.err_flag = c->err_flag.p;
    A.slice0 = (int)c->asm_first;
    return A;
}

int round_threads(int pairs) {
    int t = ((std::max(pairs, 1) + 31) / 32) * 32;
    return std::min(std::max(t, 64), 256);
}

void halo_exchange(onsas_ctx* c, double* v, int gate);
void download(onsas_ctx* c, double* h, const double* d, size_t n);
void check_deferred(onsas_ctx* c);
void p2p_wire(onsas_ctx* c, const std::vector<unsigned char*>& win, const int64_t* remote_halo_off);
void derive_element_kinds(onsas_ctx* c);
void drop_host_graph(onsas_ctx* c);
// group.inc: the same entry points on a multi-device context (global vectors in the caller's numbering)
void grp_destroy(onsas_ctx* g);
void grp_set_option(onsas_ctx* g, int32_t key, int64_t value);
void grp_materials_changed(onsas_ctx* g);
void grp_free_dofs_changed(onsas_ctx* g);
void grp_finalize(onsas_ctx* g);
void grp_wrap_single(onsas_ctx* g);
void grp_set_vec(onsas_ctx* g, int which, const double* v);
void grp_get_vec(onsas_ctx* g, int which, double* v);
void grp_add_face_load(onsas_ctx* g, int64_t n_faces, const int32_t* tri, int32_t kind, const double* values, int32_t* pattern_id);
void grp_add_nodal_load(onsas_ctx* g, int64_t n, const int32_t* nodes, const double* values, int32_t* pattern_id);
void grp_apply_loads(onsas_ctx* g, int32_t n_factors, const double* factors);
void grp_clear_loads(onsas_ctx* g);
void grp_assemble(onsas_ctx* g);
void grp_assemble_host(onsas_ctx* g, const double* U, double* F);
void grp_eval_elements(onsas_ctx* g, int32_t family, int64_t first, int64_t count, double* f, double* K, double* sig, double* eps);
void grp_step(onsas_ctx* g, bool assemble, int32_t precond, double reltol, double abstol, int64_t maxiter, int update_U, onsas_step_info* info);
void grp_pcg(onsas_ctx* g, const double* b, double* x, int32_t precond, double reltol, double abstol, int64_t maxiter, int64_t* iters, double* residual);
void grp_spmv(onsas_ctx* g, const double* x, double* y);
void grp_spmv_resident(onsas_ctx* g);
void grp_synchronize(onsas_ctx* g);
void grp_get_csr_size(onsas_ctx* g, int64_t* n_rows, int64_t* nnz);
void grp_get_csr(onsas_ctx* g, int64_t* rowptr, int32_t* col, double* val);
void grp_get_stress_strain(onsas_ctx* g, int32_t family, double* sig, double* eps);
void grp_get_table_stats(onsas_ctx* g, int64_t out[8]);
enum : int { VEC_U = 0, VEC_FEXT = 1, VEC_FINT = 2, VEC_DU = 3 };

// the assembly kernels of every element family over the slice range [c->asm_first, c->asm_first + c->asm_count)
void launch_assemble_range(onsas_ctx* c) {
    bool wrote = false;
    if (c->n_tets > 0) {
        AsmArgs A = make_asm_args(c, 0);
        const int mp = c->tab.fam[0].max_pairs_per_slice;
        const int threads = round_threads(mp);
        const size_t smem = asm_smem_bytes(A.max_pairs, A.max_width, A.max_snodes, TET_REC, 4, 3);
        switch (c->tet_kind) {
            case MAT_SVK: launch_asm_tets_kind<MAT_SVK>(c, A, threads, smem); break;
            case MAT_NEOHOOKEAN: launch_asm_tets_kind<MAT_NEOHOOKEAN>(c, A, threads, smem); break;
            case MAT_ISOLINEAR: launch_asm_tets_kind<MAT_ISOLINEAR>(c, A, threads, smem); break;
            default: launch_asm_tets_kind<MAT_MIXED>(c, A, threads, smem); break;
        }
        wrote = true;
    }
    if (c->n_trusses > 0) {
        AsmArgs A = make_asm_args(c, 1);
        const int mp = c->tab.fam[1].max_pairs_per_slice;
        const int threads = round_threads(mp);
        const size_t smem = asm_smem_bytes(A.max_pairs, A.max_width, A.max_snodes, truss_rec(c->dim), 2, c->dim);
        // register budget of the truss kernel (its CTAs are small: 128 threads on the braced lattice): 2 -> 120 registers,
        // 3 -> 80, 4 -> 64 (ONSAS_OPT_TRUSS_MINBLOCKS)
        auto go = [&](auto dimtag, auto acctag) {
            constexpr int D = decltype(dimtag)::value;
            constexpr bool ACC = decltype(acctag)::value != 0;
            if (c->truss_minb >= 4) launch_asm_inst<1, 0, D, ACC, 256, 4>(c, A, threads, smem);
            else if (c->truss_minb == 3) launch_asm_inst<1, 0, D, ACC, 256, 3>(c, A, threads, smem);
            else launch_asm_inst<1, 0, D, ACC, 256, 2>(c, A, threads, smem);
        };
        auto go_dim = [&](auto acctag) {
            if (c->dim == 3) go(IC<3>(), acctag);
            else if (c->dim == 2) go(IC<2>(), acctag);
            else go(IC<1>(), acctag);
        };
        if (wrote) go_dim(IC<1>());
        else go_dim(IC<0>());
        wrote = true;
    }
    if (!wrote) {  // structure without elements: K = 0, F_int = 0
        c->val.zero(c->stream);
        c->Fint.zero(c->stream);
    }
}

void launch_assemble(onsas_ctx* c) {
    require(c->finalized, ONSAS_ERR_NOT_READY, "onsas_finalize_mesh has not been called");
    c->co.fresh = false;  // K is about to change: the coarse inverse of the two-level preconditioner is stale
    // Multi-GPU: with the peer-memory solver the halo part of U is always current (onsas_set_U / onsas_assemble_host bring
    // it from the host, and the solver updates it together with the owned part); only the NCCL per-phase solver leaves it
    // stale, and then it is exchanged here.
    if (c->n_ranks > 1 && !(c->p2p_ready && c->cg_mode != 1)) halo_exchange(c, c->U.p, 0);
    c->asm_first = 0;
    c->asm_count = -1;
    c->asm_stream = nullptr;
    launch_assemble_range(c);
}

// ---------------------------------------------------------------- assemble! with host buffers on both sides
// U travels in and F_int travels out in pieces that overlap the kernel: the slices are cut into `host_chunks` ranges;
// range k starts as soon as the nodes its elements touch, [0, node_hi[k]), have arrived (copy stream, ascending
// prefixes of U), and its rows of F_int leave on a third stream while range k + 1 computes.  With a banded node
// numbering (any mesh numbered for locality) node_hi grows with k and the copies hide behind the kernels; with an
// arbitrary numbering node_hi[0] = n_nodes and the call degenerates to copy-in, compute, overlapped copy-out.
// K, F_int and the element records are bitwise what onsas_set_U + onsas_assemble produce (same kernel, same slices).
void build_host_plan(onsas_ctx* c) {
    auto& H = c->hp;
    const int64_t ns = c->tab.n_slices;
    const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(1, c->host_chunks), ns));
    if (H.built && (int)H.node_hi.size() == nch) return;
    drop_host_graph(c);
    host_range_plan(c->tab, c->host_chunks, c->host_mid_weight, H.slice0, H.node_hi);  // tables.cpp (CPU-tested)
    if (!H.s_in) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&H.s_in, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&H.s_out, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&H.s_k2, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&H.ev_start, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&H.ev_join[0], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&H.ev_join[1], cudaEventDisableTiming));
    }
    while ((int)H.ev_in.size() < nch) {
        cudaEvent_t a, b;
        CUDA_CHECK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        H.ev_in.push_back(a);
        H.ev_k.push_back(b);
    }
    H.built = true;
}

void drop_host_graph(onsas_ctx* c) {
    auto& H = c->hp;
    if (H.gexec) cudaGraphExecDestroy(H.gexec);
    H.gexec = nullptr;
    H.gU = H.gF = nullptr;
}

// enqueues the pipelined assembly (no host synchronisation): U in as growing prefixes on the copy-in stream, the slice ranges
// alternating on two compute streams, the rows of F_int out on the copy-out stream, the deferred-error flag last
void enqueue_host_pipeline(onsas_ctx* c, const double* U, double* F) {
    auto& H = c->hp;
    struct Restore {  // the range / stream of "the next assembly launch" never outlives this call, whatever throws
        onsas_ctx* c;
        ~Restore() {
            c->asm_first = 0;
            c->asm_count = -1;
            c->asm_stream = nullptr;
        }
    } restore{c};
    const int nch = (int)H.node_hi.size();
    const int bs = c->dim;
    CUDA_CHECK(cudaEventRecord(H.ev_start, c->stream));  // earlier work on the compute stream may still read U / F_int
    CUDA_CHECK(cudaStreamWaitEvent(H.s_in, H.ev_start, 0));
    CUDA_CHECK(cudaStreamWaitEvent(H.s_out, H.ev_start, 0));
    // consecutive ranges alternate between two compute streams: the first CTAs of range k + 1 fill the SMs that the
    // last wave of range k leaves idle (the ranges write disjoint rows of K, F_int and disjoint element records)
    const bool two = c->host_streams >= 2 && nch > 1;
    if (two) CUDA_CHECK(cudaStreamWaitEvent(H.s_k2, H.ev_start, 0));
    int64_t up = 0;  // owned nodes of U already sent
    if (c->n_nodes > c->n_owned)  // multi-GPU: the halo block first (small; every range near the interface reads it)
        CUDA_CHECK(cudaMemcpyAsync(c->U.p + c->n_owned * bs, U + c->n_owned * bs, (size_t)(c->n_nodes - c->n_owned) * bs * sizeof(double), cudaMemcpyHostToDevice, H.s_in));
    for (int k = 0; k < nch; ++k) {
        cudaStream_t sk = (two && (k & 1)) ? H.s_k2 : c->stream;
        const int64_t hi = k + 1 == nch ? c->n_owned : H.node_hi[k];  // the last piece takes what no element touches
        if (hi > up) {
            CUDA_CHECK(cudaMemcpyAsync(c->U.p + up * bs, U + up * bs, (size_t)(hi - up) * bs * sizeof(double), cudaMemcpyHostToDevice, H.s_in));
            up = hi;
        }
        CUDA_CHECK(cudaEventRecord(H.ev_in[k], H.s_in));
        CUDA_CHECK(cudaStreamWaitEvent(sk, H.ev_in[k], 0));
        c->asm_first = H.slice0[k];
        c->asm_count = H.slice0[k + 1] - H.slice0[k];
        c->asm_stream = sk;
        if (c->asm_count > 0) launch_assemble_range(c);
        c->asm_stream = nullptr;
        CUDA_CHECK(cudaEventRecord(H.ev_k[k], sk));
        CUDA_CHECK(cudaStreamWaitEvent(H.s_out, H.ev_k[k], 0));
        if (sk != c->stream) CUDA_CHECK(cudaStreamWaitEvent(c->stream, H.ev_k[k], 0));  // later work on the context's stream sees every range
        const int64_t r0 = std::min<int64_t>(H.slice0[k] * SLICE_ROWS, c->n_owned), r1 = std::min<int64_t>(H.slice0[k + 1] * SLICE_ROWS, c->n_owned);
        if (r1 > r0)
            CUDA_CHECK(cudaMemcpyAsync(F + r0 * bs, c->Fint.p + r0 * bs, (size_t)(r1 - r0) * bs * sizeof(double), cudaMemcpyDeviceToHost, H.s_out));
    }
    CUDA_CHECK(cudaMemcpyAsync(c->h_flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, H.s_out));
}

void assemble_host(onsas_ctx* c, const double* U, double* F) {
    require(c->finalized, ONSAS_ERR_NOT_READY, "onsas_finalize_mesh has not been called");
    const size_t nl = (size_t)c->n_local_dofs(), no = (size_t)c->n_own_dofs();
    if (c->host_chunks <= 1 || c->tab.n_slices == 0 || (c->n_tets == 0 && c->n_trusses == 0)) {
        // pipelining switched off: copy in, assemble, copy out.  (Multi-GPU pipelines too: U arrives with its halo part.)
        CUDA_CHECK(cudaMemcpyAsync(c->U.p, U, nl * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        launch_assemble(c);
        if (nl > no) std::fill(F + no, F + nl, 0.0);
        CUDA_CHECK(cudaMemcpyAsync(c->h_flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        download(c, F, c->Fint.p, no);
        check_deferred(c);
        return;
    }
    build_host_plan(c);
    auto& H = c->hp;
    c->co.fresh = false;
    if (nl > no) std::fill(F + no, F + nl, 0.0);  // halo part of F_int: not assembled on this rank
    // The pipeline is ~80 stream operations (copies, kernels, events on four streams): enqueueing them costs the host about as
    // much time as the device needs to run them.  When the caller keeps its state in the same pinned buffers from call to call
    // (the reference's state vectors are persistent arrays), the whole pipeline is captured ONCE into a CUDA graph and every
    // later call is a single graph launch.  Same operations, same order on every stream: bitwise the same results.
    auto is_pinned = [](const void* p) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeHost;
    };
    bool done = false;
    if (c->host_graph && !H.graph_failed) {
        if (H.gexec && H.gU == U && H.gF == F) {
            CUDA_CHECK(cudaGraphLaunch(H.gexec, c->stream));
            done = true;
        } else if (H.seenU == U && H.seenF == F && is_pinned(U) && is_pinned(F)) {
            drop_host_graph(c);
            cudaGraph_t g = nullptr;
            bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                try {
                    enqueue_host_pipeline(c, U, F);
                    // join the forked streams back into the capturing one
                    CUDA_CHECK(cudaEventRecord(H.ev_join[0], H.s_in));
                    CUDA_CHECK(cudaStreamWaitEvent(c->stream, H.ev_join[0], 0));
                    CUDA_CHECK(cudaEventRecord(H.ev_join[1], H.s_out));
                    CUDA_CHECK(cudaStreamWaitEvent(c->stream, H.ev_join[1], 0));
                } catch (...) {
                    ok = false;
                }
                if (cudaStreamEndCapture(c->stream, &g) != cudaSuccess || g == nullptr) ok = false;
            }
            if (ok && cudaGraphInstantiate(&H.gexec, g, 0) != cudaSuccess) ok = false;
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            if (ok) {
                H.gU = U;
                H.gF = F;
                CUDA_CHECK(cudaGraphLaunch(H.gexec, c->stream));
                done = true;
            } else {
                H.gexec = nullptr;
                H.graph_failed = true;  // this driver / pipeline shape cannot be captured: stay with the eager path
            }
        }
    }
    H.seenU = U;
    H.seenF = F;
    if (done) {
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
        enqueue_host_pipeline(c, U, F);
        CUDA_CHECK(cudaStreamSynchronize(H.s_out));  // the last kernel has finished too: the compute stream is idle
    }
    check_deferred(c);
}

// ---------------------------------------------------------------- halo exchange (NCCL send/recv)
void halo_exchange(onsas_ctx* c, double* v, int gate) {
    if (c->n_ranks <= 1 || c->nbr_rank.empty()) return;
    require(c->comm != nullptr, ONSAS_ERR_COMM, "this multi-GPU context has no NCCL communicator (peer-memory solver only: ONSAS_OPT_CG_MODE 0 or 2)");
    const int bs = c->dim;
    const int64_t n_send = c->send_ptr.back();
    if (n_send > 0) {
        const int64_t tot = n_send * bs;
        k_pack<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(v, c->send_nodes.p, n_send, bs, c->sendbuf.p, c->st.p, gate);
        CUDA_CHECK(cudaGetLastError());
    }
    NCCL_CHECK(g_nccl.GroupStart());
    for (size_t k = 0; k < c->nbr_rank.size(); ++k) {
        const int64_t ns = (c->send_ptr[k + 1] - c->send_ptr[k]) * bs, nr = (c->recv_ptr[k + 1] - c->recv_ptr[k]) * bs;
        if (ns > 0) NCCL_CHECK(g_nccl.Send(c->sendbuf.p + c->send_ptr[k] * bs, (size_t)ns, ncclDouble, c->nbr_rank[k], c->comm, c->stream));
        if (nr > 0) NCCL_CHECK(g_nccl.Recv(v + (c->n_owned + c->recv_ptr[k]) * bs, (size_t)nr, ncclDouble, c->nbr_rank[k], c->comm, c->stream));
    }
    NCCL_CHECK(g_nccl.GroupEnd());
}

void allreduce(onsas_ctx* c, double* d, int count) {
    if (c->n_ranks <= 1) return;
    require(c->comm != nullptr, ONSAS_ERR_COMM, "this multi-GPU context has no NCCL communicator (peer-memory solver only: ONSAS_OPT_CG_MODE 0 or 2)");
    NCCL_CHECK(g_nccl.AllReduce(d, d, (size_t)count, ncclDouble, ncclSum, c->comm, c->stream));
}

// ---------------------------------------------------------------- P2P window layout: [slots (LL)][epochs][zh (LL)]
constexpr size_t P2P_SLOTS_BYTES = 2 * P2P_MAXR * 4 * 16;
constexpr size_t P2P_HDR_BYTES = 4096;
static_assert(P2P_SLOTS_BYTES + 16 <= P2P_HDR_BYTES, "window header too small");
inline unsigned long long* win_slots(unsigned char* w) { return reinterpret_cast<unsigned long long*>(w); }
inline unsigned long long* win_epochs(unsigned char* w) { return reinterpret_cast<unsigned long long*>(w + P2P_SLOTS_BYTES); }
// global coarse level (same offsets on every rank): [E2: NC2^2 doubles][w2 LL: 2 x NC2 pairs][row flags: 16 words][w2 epoch, flag]
constexpr size_t P2P_E2_BYTES = (size_t)COARSE_NC_MAX * COARSE_NC_MAX * 8;
constexpr size_t P2P_W2_BYTES = (size_t)2 * COARSE_NC_MAX * 16;
constexpr size_t P2P_COARSE_BYTES = P2P_E2_BYTES + P2P_W2_BYTES + 256;
inline double* win_E2(unsigned char* w) { return reinterpret_cast<double*>(w + P2P_HDR_BYTES); }
inline unsigned long long* win_w2(unsigned char* w) { return reinterpret_cast<unsigned long long*>(w + P2P_HDR_BYTES + P2P_E2_BYTES); }
inline unsigned long long* win_rowflags(unsigned char* w) { return reinterpret_cast<unsigned long long*>(w + P2P_HDR_BYTES + P2P_E2_BYTES + P2P_W2_BYTES); }
inline unsigned long long* win_w2epoch(unsigned char* w) { return win_rowflags(w) + 16; }
inline unsigned int* win_w2flag(unsigned char* w) { return reinterpret_cast<unsigned int*>(win_rowflags(w) + 17); }
inline unsigned long long* win_zh(unsigned char* w) { return reinterpret_cast<unsigned long long*>(w + P2P_HDR_BYTES + P2P_COARSE_BYTES); }

// ---------------------------------------------------------------- CG drivers
CgArgs make_cg_args(onsas_ctx* c, int precond, double reltol, double abstol, int64_t maxiter, bool use_rhs, int update_U) {
    CgArgs A{};
    A.n_rows = c->n_owned;
    A.n = c->n_own_dofs();
    A.slice_ptr = c->slice_ptr.p;
    A.col = c->col.p;
    A.val = c->val.p;
    A.diag_slot = c->diag_slot.p;
    A.mask = c->mask.p;
    A.x = c->x.p;
    A.r = c->r.p;
    A.p = c->p.p;
    A.Ap = c->Ap.p;
    A.dinv = c->dinv.p;
    A.Fext = c->Fext.p;
    A.Fint = c->Fint.p;
    A.rhs = use_rhs ? c->rhs.p : nullptr;
    A.U = c->U.p;
    A.update_U = update_U;
    A.precond = precond;
    A.reltol = reltol;
    A.abstol = abstol;
    A.maxiter = maxiter > 0 ? maxiter : c->n_free_global;
    A.partials = c->partials.p;
    A.part_stride = c->part_stride;
    A.st = c->st.p;
    A.prof = c->cg_profile ? c->prof.p : nullptr;
    A.err = c->err_flag.p;
    A.p_pad = c->p_pad.p;
    A.s = c->s_vec.p;
    return A;
}

// grid of the stand-alone SpMV: one thread per scalar row of the padded slices, capped to the partial-sum stride
int spmv_grid(onsas_ctx* c) {
    const int64_t items = c->tab.n_slices * (int64_t)(SLICE_ROWS * c->dim);
    return (int)std::max<int64_t>(1, std::min<int64_t>((items + CG_THREADS - 1) / CG_THREADS, (int64_t)c->part_stride));
}

// The persistent CG kernel is compiled for 4, 5 or 6 resident CTAs per SM (64 / 48 / 40 registers);
// ONSAS_OPT_CG_BLOCKS_PER_SM picks the variant (default 4).
template <int BS>
void* persistent_kernel(onsas_ctx* c) {
    const int v = c->cg_bps >= 6 ? 6 : c->cg_bps == 5 ? 5 : 4;
    if (c->cg_profile) return v == 6 ? (void*)cg_persistent<BS, true, 6> : v == 5 ? (void*)cg_persistent<BS, true, 5> : (void*)cg_persistent<BS, true, 4>;
    return v == 6 ? (void*)cg_persistent<BS, false, 6> : v == 5 ? (void*)cg_persistent<BS, false, 5> : (void*)cg_persistent<BS, false, 4>;
}

template <int BS>
int persistent_grid(onsas_ctx* c) {
    int bps = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, persistent_kernel<BS>(c), CG_THREADS, 0));
    require(bps > 0, ONSAS_ERR_CUDA, "persistent CG kernel does not fit on an SM");
    if (c->cg_bps > 0 && c->cg_bps < 4) bps = std::min(bps, c->cg_bps);
    int g = bps * c->n_sm;
    return std::min(g, c->part_stride);
}

// ---------------------------------------------------------------- streamed persistent CG (cg_stream): plan + launch
P2PArgs make_p2p_args(onsas_ctx* c) {
    P2PArgs P{};  // n_ranks = 0: single GPU
    if ((c->n_ranks > 1 || c->force_mg) && c->p2p_ready) {
        P.n_ranks = c->n_ranks;
        P.rank = c->rank;
        P.n_halo_dofs = (long long)(c->n_local_dofs() - c->n_own_dofs());
        P.push_ptr = c->d_push_ptr.p;
        P.push_dst = c->d_push_dst.p;
        P.zh = win_zh(c->window.p);
        P.slots = win_slots(c->window.p);
        P.peer_slots = c->d_peer_slots.p;
        P.epochs = win_epochs(c->window.p);
        P.err = c->err_flag.p;
    }
    return P;
}

// Ring geometry of cg_stream: a slot holds the widest slice.  Preferred: 12 consumer warps x 2 slots (one slice per
// warp in flight while it computes on the other); fewer warps for wide rows.  Returns false when not even
// one warp with two slots fits: the register-fed kernel runs then.
template <int BS>
bool plan_stream(onsas_ctx* c) {
    if (c->st_plan.built) return c->st_plan.ok;
    c->st_plan.built = true;
    c->st_plan.ok = false;
    if (c->tab.n_slices == 0 || c->tab.max_width <= 0) return false;
    int optin = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    const size_t budget = (size_t)optin - 3072;  // static shared memory of the kernel + slack
    const size_t slot_bytes = (size_t)(BS * BS * 8 + 4) * SLICE_ROWS * (size_t)c->tab.max_width;
    const int slots = (int)(budget / slot_bytes);
    if (slots < 2) return false;
    // measured on the 1 M-tet cube (scripts/cg_stream_probe.py): 12 warps x 2 slots 47.8 us / CG iteration, 8 x 3 52.8 us
    const int cw_max = 12;
    int want_depth = 2;
    if (const char* e = getenv("ONSAS_STREAM_DEPTH")) want_depth = std::max(2, std::min(ST_MAX_DEPTH, atoi(e)));  // experiment knob
    int depth = std::min(want_depth, slots);
    int n_cw = std::min(cw_max, slots / depth);
    if (n_cw < cw_max && depth > 2) {  // wide rows: rather keep the warps and give each two slots
        depth = 2;
        n_cw = std::min(cw_max, slots / 2);
    }
    const size_t smem = slot_bytes * (size_t)n_cw * depth;
    void* kern = (void*)cg_stream<BS, 12, false>;
    void* kern_sr = (void*)cg_stream<BS, 12, true>;
    const int threads = (cw_max + 1) * 32;
    ensure_dyn_smem(c->device, kern, smem);
    ensure_dyn_smem(c->device, kern_sr, smem);
    int occ = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return false;
    c->st_plan.kern = kern;
    c->st_plan.kern_sr = kern_sr;
    c->st_plan.threads = threads;
    c->st_plan.n_cw = n_cw;
    c->st_plan.depth = depth;
    c->st_plan.smem = smem;
    c->st_plan.grid = std::min(c->n_sm, c->part_stride);
    c->st_plan.ok = true;
    if (getenv("ONSAS_VERBOSE"))
        fprintf(stderr, "[onsas] streamed CG: %d CTAs x (%d consumer warps + 1 producer warp), %d slots of %.1f KB per warp, %.1f KB of shared memory\n",
                c->st_plan.grid, n_cw, depth, slot_bytes / 1024.0, smem / 1024.0);
    return true;
}

template <int BS>
void launch_stream(onsas_ctx* c, CgArgs A) {
    StreamArgs S{};
    S.n_cw = c->st_plan.n_cw;
    S.depth = c->st_plan.depth;
    S.slot_blocks = c->tab.max_width;
    S.n_slices = c->tab.n_slices;
    S.l2_prefetch = c->cg_l2_prefetch;
    if (const char* e = getenv("ONSAS_STREAM_L2_PREFETCH")) S.l2_prefetch = std::max(0, std::min(64, atoi(e)));  // experiment knob
    P2PArgs P = make_p2p_args(c);
    void* args[] = {&A, &S, &P};
    // Jacobi-PCG (the north-star solver) runs the single-reduction recurrence; precond = 0 keeps the classic one, which
    // restates IterativeSolvers' cg! step by step, and the two-level preconditioner needs its own phase structure
    const bool sr = (A.precond == 1 && (c->cg_single_reduction & 1)) || (A.precond == 2 && (c->cg_single_reduction & 2) && !c->co.glob);
    CUDA_CHECK(cudaLaunchCooperativeKernel(sr ? c->st_plan.kern_sr : c->st_plan.kern, dim3(c->st_plan.grid), dim3(c->st_plan.threads), args, c->st_plan.smem, c->stream));
}

// ---------------------------------------------------------------- two-level preconditioner: aggregates + coarse inverse
constexpr int CO_NC_MAX = COARSE_NC_MAX;  // partition.hpp (the partitioner cuts the same aggregates when it numbers aggregate-major)
constexpr int CO_TARGET_NODES = COARSE_TARGET_NODES;

// k-way recursive coordinate bisection of the owned nodes (deterministic: ties broken by node id)
void rcb_aggregate(const double* xyz, int dim, std::vector<int32_t>& ids, size_t lo, size_t hi, int parts, int first, std::vector<int32_t>& agg) {
    if (parts <= 1 || hi - lo <= 1) {
        for (size_t k = lo; k < hi; ++k) agg[ids[k]] = first;
        return;
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (size_t k = lo; k < hi; ++k)
        for (int d = 0; d < dim; ++d) {
            const double v = xyz[(size_t)ids[k] * dim + d];
            mn[d] = std::min(mn[d], v);
            mx[d] = std::max(mx[d], v);
        }
    int ax = 0;
    for (int d = 1; d < dim; ++d)
        if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
    const int p1 = parts / 2;
    const size_t mid = lo + (hi - lo) * (size_t)p1 / (size_t)parts;
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int32_t a, int32_t b) {
        const double va = xyz[(size_t)a * dim + ax], vb = xyz[(size_t)b * dim + ax];
        return va < vb || (va == vb && a < b);
    });
    rcb_aggregate(xyz, dim, ids, lo, mid, p1, first, agg);
    rcb_aggregate(xyz, dim, ids, mid, hi, parts - p1, first + p1, agg);
}

void build_coarse(onsas_ctx* c) {
    if (c->co.built) return;
    const int64_t n = c->n_owned;
    require(n > 0, ONSAS_ERR_NOT_READY, "two-level preconditioner: no owned nodes");
    const int cd = (c->dim == 3 && c->coarse_rbm) ? 6 : c->dim;
    int n_agg = (int)std::max<int64_t>(1, std::min<int64_t>(n / CO_TARGET_NODES, CO_NC_MAX / cd));
    std::vector<int32_t> ids((size_t)n), agg((size_t)n, 0);
    const bool glob = c->coarse_global && c->n_ranks > 1 && c->p2p_ready && cd == 6 && c->n_agg2_per_rank > 0 &&
                      (int64_t)c->h_agg2.size() == c->n_nodes && c->cg_mode == 0;
    c->co.glob = glob;
    std::vector<int32_t> child_ptr, parent_of;
    if (glob) {
        // the rank's aggregates are cut INSIDE its level-2 aggregates (Z2 = Z T: the global level folds into the rank's operator)
        const int n2 = c->n_agg2_per_rank, first2 = c->rank * n2;
        const int per = std::max(1, n_agg / n2);
        std::vector<std::vector<int32_t>> lists((size_t)n2);
        for (int64_t i = 0; i < n; ++i) {
            const int P2 = c->h_agg2[(size_t)i] - first2;
            require(P2 >= 0 && P2 < n2, ONSAS_ERR_INVALID_ARG, "level-2 aggregate of an owned node belongs to another rank");
            lists[(size_t)P2].push_back((int32_t)i);
        }
        child_ptr.assign(1, 0);
        int next = 0;
        for (int P2 = 0; P2 < n2; ++P2) {
            std::vector<int32_t>& L = lists[(size_t)P2];
            const int k = (int)std::max<size_t>(1, std::min<size_t>((size_t)per, L.size()));
            if (!L.empty()) rcb_aggregate(c->h_xyz.data(), c->dim, L, 0, L.size(), k, next, agg);
            for (int j = 0; j < k; ++j) parent_of.push_back(P2);
            next += k;
            child_ptr.push_back(next);
        }
        n_agg = next;
        c->co.n2_own = n2;
        c->co.agg2_first = first2;
        c->co.nc2 = (int)(c->h_cen2.size() / 3) * 6;
    } else if (!c->h_agg_ptr.empty() && c->h_agg_ptr.back() == n && (int)(c->h_agg_ptr.size() - 1) * cd <= CO_NC_MAX) {
        // aggregate-major numbering (ONSAS_OPT_REORDER = 2): the partitioner already cut the aggregates, they are node ranges
        n_agg = (int)c->h_agg_ptr.size() - 1;
        for (int a = 0; a < n_agg; ++a)
            for (int32_t i = c->h_agg_ptr[a]; i < c->h_agg_ptr[a + 1]; ++i) agg[i] = a;
    } else {
        for (int64_t i = 0; i < n; ++i) ids[i] = (int32_t)i;
        rcb_aggregate(c->h_xyz.data(), c->dim, ids, 0, (size_t)n, n_agg, 0, agg);
    }
    std::vector<int32_t> ptr((size_t)n_agg + 1, 0), nodes((size_t)n);
    for (int64_t i = 0; i < n; ++i) ptr[agg[i] + 1]++;
    for (int a = 0; a < n_agg; ++a) ptr[a + 1] += ptr[a];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) nodes[fill[agg[i]]++] = (int32_t)i;  // ascending node id inside every aggregate
    if (cd == 6) {  // node positions relative to the aggregate's centroid (reference configuration): the rotation modes
        std::vector<double> cen((size_t)n_agg * 3, 0.0), rho((size_t)n * 3);
        for (int a = 0; a < n_agg; ++a) {
            for (int32_t q = ptr[a]; q < ptr[a + 1]; ++q)
                for (int d = 0; d < 3; ++d) cen[(size_t)a * 3 + d] += c->h_xyz[(size_t)nodes[q] * 3 + d];
            const double inv = 1.0 / std::max<int32_t>(1, ptr[a + 1] - ptr[a]);
            for (int d = 0; d < 3; ++d) cen[(size_t)a * 3 + d] *= inv;
        }
        for (int64_t i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) rho[(size_t)i * 3 + d] = c->h_xyz[(size_t)i * 3 + d] - cen[(size_t)agg[i] * 3 + d];
        c->co_rho.upload(rho, c->stream);
    }
    const int nc = n_agg * cd;
    cudaStream_t s = c->stream;
    if (glob) {
        // tables of the global level: level-2 aggregate and position relative to ITS centroid for every local node (halo
        // nodes too: the rows of E2 couple to the neighbours' aggregates), the node lists of the own level-2 aggregates,
        // parent and centroid offset of every own aggregate
        const int n2 = c->co.n2_own, first2 = c->co.agg2_first;
        std::vector<double> rho2((size_t)c->n_nodes * 3), dvec((size_t)n_agg * 3), cen1((size_t)n_agg * 3, 0.0);
        for (int64_t i = 0; i < c->n_nodes; ++i)
            for (int d = 0; d < 3; ++d) rho2[(size_t)i * 3 + d] = c->h_xyz[(size_t)i * 3 + d] - c->h_cen2[(size_t)c->h_agg2[(size_t)i] * 3 + d];
        for (int a = 0; a < n_agg; ++a) {
            for (int32_t q = ptr[a]; q < ptr[a + 1]; ++q)
                for (int d = 0; d < 3; ++d) cen1[(size_t)a * 3 + d] += c->h_xyz[(size_t)nodes[q] * 3 + d];
            for (int d = 0; d < 3; ++d) {
                cen1[(size_t)a * 3 + d] /= std::max<int32_t>(1, ptr[a + 1] - ptr[a]);
                dvec[(size_t)a * 3 + d] = cen1[(size_t)a * 3 + d] - c->h_cen2[(size_t)(first2 + parent_of[(size_t)a]) * 3 + d];
            }
        }
        std::vector<int32_t> ptr2((size_t)n2 + 1, 0), nodes2((size_t)n);
        for (int64_t i = 0; i < n; ++i) ptr2[(size_t)(c->h_agg2[(size_t)i] - first2) + 1]++;
        for (int a = 0; a < n2; ++a) ptr2[(size_t)a + 1] += ptr2[(size_t)a];
        std::vector<int32_t> fill2(ptr2.begin(), ptr2.end() - 1);
        for (int64_t i = 0; i < n; ++i) nodes2[(size_t)fill2[(size_t)(c->h_agg2[(size_t)i] - first2)]++] = (int32_t)i;
        c->co2_agg.upload(c->h_agg2, s);
        c->co2_rho.upload(rho2, s);
        c->co2_ptr.upload(ptr2, s);
        c->co2_nodes.upload(nodes2, s);
        c->co_parent.upload(parent_of, s);
        c->co_child_ptr.upload(child_ptr, s);
        c->co_dvec.upload(dvec, s);
        c->co_G.alloc((size_t)nc * c->co.nc2);
        c->co_w2plain.alloc((size_t)c->co.nc2);
        c->co_rowbuf2.alloc((size_t)2 * GJ_B * c->co.nc2);
    }
    c->co_agg.upload(agg, s);
    c->co_agg_ptr.upload(ptr, s);
    c->co_agg_nodes.upload(nodes, s);
    c->co_E.alloc((size_t)nc * nc);
    c->co_w.alloc((size_t)nc);
    c->co_y.alloc((size_t)nc);
    c->co_rowbuf.alloc((size_t)2 * GJ_B * nc);
    CUDA_CHECK(cudaStreamSynchronize(s));
    c->co.n_agg = n_agg;
    c->co.nc = nc;
    c->co.cd = cd;
    c->co.built = true;
    c->co.fresh = false;
    if (getenv("ONSAS_VERBOSE")) fprintf(stderr, "[onsas] two-level preconditioner: %d aggregates of ~%lld nodes, %d coarse dofs each (%s), %d in total\n", n_agg, (long long)(n / n_agg), cd, cd == 6 ? "translations + rotations" : "translations", nc);
}

void fill_coarse_args(onsas_ctx* c, CgArgs& A) {
    A.co.n_agg = c->co.n_agg;
    A.co.nc = c->co.nc;
    A.co.cd = c->co.cd;
    A.co.fused = c->coarse_fused ? 1 : 0;
    A.co.rho = c->co.cd == 6 ? c->co_rho.p : nullptr;
    A.co.agg = c->co_agg.p;
    A.co.agg_ptr = c->co_agg_ptr.p;
    A.co.agg_nodes = c->co_agg_nodes.p;
    A.co.Einv = c->co_E.p;
    A.co.w = c->co_w.p;
    A.co.y = c->co_y.p;
    A.co.n_cols = (int)c->n_owned;
    A.co.agg_row0 = 0;
    A.co.glob = c->co.glob ? 1 : 0;
    A.co.n_ranks = c->n_ranks;
    if (c->co.glob) {
        A.co.nc2 = c->co.nc2;
        A.co.n2_own = c->co.n2_own;
        A.co.agg2_first = c->co.agg2_first;
        A.co.G = c->co_G.p;
        A.co.child_ptr = c->co_child_ptr.p;
        A.co.dvec = c->co_dvec.p;
        A.co.w2_ll = win_w2(c->window.p);
        A.co.peer_w2 = c->d_peer_w2.p;
        A.co.w2_plain = c->co_w2plain.p;
        A.co.w2_flag = win_w2flag(c->window.p);
        A.co.w2_epoch = win_w2epoch(c->window.p);
    }
}

// E = Z^T (M K M) Z for the K currently in memory, then its explicit inverse (both deterministic)
template <int BS>
void refresh_coarse(onsas_ctx* c, CgArgs& A) {
    build_coarse(c);
    fill_coarse_args(c, A);
    if (c->co.fresh) return;
    const int nc = c->co.nc;
    {
        const size_t smem = (size_t)c->co.cd * nc * sizeof(double);
        static const bool serial = getenv("ONSAS_COARSE_SERIAL") != nullptr;  // experiment knob: the first form of the kernel
        ensure_dyn_smem(c->device, (const void*)k_coarse_assemble<BS>, smem);
        ensure_dyn_smem(c->device, (const void*)k_coarse_assemble_serial<BS>, smem);
        if (serial) k_coarse_assemble_serial<BS><<<c->co.n_agg, CO_THREADS, smem, c->stream>>>(A, c->co_E.p);
        else k_coarse_assemble<BS><<<c->co.n_agg, CO_THREADS, smem, c->stream>>>(A, c->co_E.p);
        CUDA_CHECK(cudaGetLastError());
        if (getenv("ONSAS_COARSE_CHECK")) {  // diagnostics: both forms of the kernel on the same K, largest difference of E
            DevBuf<double> E2;
            E2.alloc((size_t)nc * nc);
            if (serial) k_coarse_assemble<BS><<<c->co.n_agg, CO_THREADS, smem, c->stream>>>(A, E2.p);
            else k_coarse_assemble_serial<BS><<<c->co.n_agg, CO_THREADS, smem, c->stream>>>(A, E2.p);
            CUDA_CHECK(cudaGetLastError());
            std::vector<double> h1((size_t)nc * nc), h2((size_t)nc * nc);
            CUDA_CHECK(cudaMemcpyAsync(h1.data(), c->co_E.p, h1.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(h2.data(), E2.p, h2.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            double dmax = 0.0, emax = 0.0, asym = 0.0;
            for (size_t k = 0; k < h1.size(); ++k) {
                dmax = std::max(dmax, std::fabs(h1[k] - h2[k]));
                emax = std::max(emax, std::fabs(h2[k]));
            }
            for (int i = 0; i < nc; ++i)
                for (int j = 0; j < i; ++j) asym = std::max(asym, std::fabs(h1[(size_t)i * nc + j] - h1[(size_t)j * nc + i]));
            fprintf(stderr, "[onsas] coarse operator check: nc = %d, max |E - E_other_form| = %.3e, max |E| = %.3e, max |E - E^T| = %.3e\n", nc, dmax, emax, asym);
        }
    }
    if (c->gj_blocked && (nc + GJ_B - 1) / GJ_B <= c->n_sm) {  // panels of 12 rows: nc / 12 grid barriers
        const int grid = (nc + GJ_B - 1) / GJ_B;
        const size_t smem = ((size_t)GJ_B * nc + 2 * GJ_B * GJ_B) * sizeof(double);
        ensure_dyn_smem(c->device, (const void*)k_gj_invert_blocked, smem);
        double* M = c->co_E.p;
        int ncv = nc;
        double* rb = c->co_rowbuf.p;
        void* args[] = {&M, &ncv, &rb};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_gj_invert_blocked, dim3(grid), dim3(GJ_THREADS), args, smem, c->stream));
    } else {
        const int rows = (nc + c->n_sm - 1) / c->n_sm;
        require(rows <= GJ_MAX_ROWS, ONSAS_ERR_UNSUPPORTED, "coarse space too large for the Gauss-Jordan kernel");
        const int grid = (nc + rows - 1) / rows;
        const size_t smem = (size_t)rows * nc * sizeof(double);
        ensure_dyn_smem(c->device, (const void*)k_gj_invert, smem);
        double* M = c->co_E.p;
        int ncv = nc, rowsv = rows;
        double* rb = c->co_rowbuf.p;
        void* args[] = {&M, &ncv, &rowsv, &rb};
        CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_gj_invert, dim3(grid), dim3(GJ_THREADS), args, smem, c->stream));
    }
    if constexpr (BS == 3) {
        if (c->co.glob) {
            // ---- global level: this rank's rows of E2 = Z2^T (M K M) Z2 (halo columns included), pushed into every rank's copy
            //      of E2 over the peer window; once everybody's rows have landed each rank inverts E2 and folds it into
            //      G = T (E2^-1)[own rows, :].  Refreshes of different ranks cannot overlap: a globally synchronised solve
            //      separates them.
            const int nc2 = c->co.nc2, n2 = c->co.n2_own, first2 = c->co.agg2_first;
            require((nc2 + GJ_B - 1) / GJ_B <= c->n_sm && nc2 <= COARSE_NC_MAX, ONSAS_ERR_UNSUPPORTED, "global coarse space too large");
            CgArgs A2 = A;
            A2.co.n_agg = n2;
            A2.co.nc = nc2;
            A2.co.cd = 6;
            A2.co.rho = c->co2_rho.p;
            A2.co.agg = c->co2_agg.p;
            A2.co.agg_ptr = c->co2_ptr.p;
            A2.co.agg_nodes = c->co2_nodes.p;
            A2.co.n_cols = (int)c->n_nodes;
            A2.co.agg_row0 = first2;
            double* E2 = win_E2(c->window.p);
            double* rows = E2 + (size_t)first2 * 6 * nc2;
            const size_t smem2 = (size_t)6 * nc2 * sizeof(double);
            ensure_dyn_smem(c->device, (const void*)k_coarse_assemble<BS>, smem2);
            k_coarse_assemble<BS><<<n2, CO_THREADS, smem2, c->stream>>>(A2, rows);
            CUDA_CHECK(cudaGetLastError());
            const size_t nrow = (size_t)n2 * 6 * nc2;
            k_push_rows<<<64, 256, 0, c->stream>>>(rows, nrow, c->d_peer_E2.p, c->n_ranks, c->rank, (size_t)first2 * 6 * nc2);
            ++c->co2_epoch;
            k_set_flags<<<1, 32, 0, c->stream>>>(c->d_peer_rowflags.p, c->n_ranks, c->rank, c->co2_epoch);
            k_wait_flags<<<1, 32, 0, c->stream>>>(win_rowflags(c->window.p), c->n_ranks, c->rank, c->co2_epoch, c->err_flag.p);
            CUDA_CHECK(cudaGetLastError());
            {
                const int grid = (nc2 + GJ_B - 1) / GJ_B;
                const size_t smem = ((size_t)GJ_B * nc2 + 2 * GJ_B * GJ_B) * sizeof(double);
                ensure_dyn_smem(c->device, (const void*)k_gj_invert_blocked, smem);
                double* M = E2;
                int ncv = nc2;
                double* rb = c->co_rowbuf2.p;
                void* args[] = {&M, &ncv, &rb};
                CUDA_CHECK(cudaLaunchCooperativeKernel((void*)k_gj_invert_blocked, dim3(grid), dim3(GJ_THREADS), args, smem, c->stream));
            }
            k_build_G<<<nc, 256, 0, c->stream>>>(E2, nc2, nc, c->co_parent.p, first2, c->co_dvec.p, c->co_G.p);
            CUDA_CHECK(cudaGetLastError());
        }
    }
    c->co.fresh = true;
}

template <int BS>
void run_cg_bs(onsas_ctx* c, CgArgs A) {
    const int64_t n = A.n;
    // persistent solver: single GPU, or N GPUs once the peer-memory window is imported; without the window a
    // multi-rank solve runs the multi-launch driver below (NCCL between the phases)
    if (A.precond == 2) {
        require(c->cg_mode == 0 && (c->n_ranks == 1 || c->p2p_ready) && plan_stream<BS>(c), ONSAS_ERR_UNSUPPORTED,
                "the two-level preconditioner runs in the streamed persistent solver only (ONSAS_OPT_CG_MODE = 0)");
        refresh_coarse<BS>(c, A);
    }
    if (c->cg_mode != 1 && (c->n_ranks == 1 || c->p2p_ready)) {
        if (c->cg_mode == 0 && plan_stream<BS>(c)) {
            launch_stream<BS>(c, A);
            return;
        }
        if (c->cg_grid == 0) c->cg_grid = persistent_grid<BS>(c);
        P2PArgs P = make_p2p_args(c);
        void* args[] = {&A, &P};
        CUDA_CHECK(cudaLaunchCooperativeKernel(persistent_kernel<BS>(c), dim3(c->cg_grid), dim3(CG_THREADS), args, 0, c->stream));
        return;
    }
    // one launch per phase; collectives in-stream between them
    const int G = (int)std::max<int64_t>(1, std::min<int64_t>((n + CG_THREADS - 1) / CG_THREADS, (int64_t)c->n_sm * 8));
    const int Gr = spmv_grid(c);
    cudaStream_t s = c->stream;
    double* red = c->red.p;
    k_cg_prologue<BS><<<G, CG_THREADS, 0, s>>>(A);
    k_reduce_partials<<<1, CG_THREADS, 0, s>>>(A.partials, A.part_stride, G, 0, 4, red, A.st, 0);
    allreduce(c, red, 4);
    k_cg_init_state<<<1, 1, 0, s>>>(A, red);
    CUDA_CHECK(cudaGetLastError());
    for (;;) {
        for (int j = 0; j < c->check_every; ++j) {
            k_cg_update_p<<<G, CG_THREADS, 0, s>>>(A);
            halo_exchange(c, A.p, 1);
            k_spmv_dot<BS><<<Gr, CG_THREADS, 0, s>>>(A, 1);
            k_reduce_partials<<<1, CG_THREADS, 0, s>>>(A.partials, A.part_stride, Gr, P_PAP, 1, red, A.st, 1);
            allreduce(c, red + P_PAP, 1);
            k_cg_update_xr<<<G, CG_THREADS, 0, s>>>(A, red);
            k_reduce_partials<<<1, CG_THREADS, 0, s>>>(A.partials, A.part_stride, G, 0, 2, red, A.st, 1);
            allreduce(c, red, 2);
            k_cg_advance<<<1, 1, 0, s>>>(A, red);
        }
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(c->h_flag, &A.st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK(cudaStreamSynchronize(s));
        if (*c->h_flag) break;
    }
    k_cg_epilogue<<<G, CG_THREADS, 0, s>>>(A);
    k_reduce_partials<<<1, CG_THREADS, 0, s>>>(A.partials, A.part_stride, G, P_DD, 1, red, A.st, 0);
    allreduce(c, red + P_DD, 1);
    k_cg_finish<<<1, 1, 0, s>>>(A, red);
    CUDA_CHECK(cudaGetLastError());
}

void run_cg(onsas_ctx* c, const CgArgs& A) {
    require(c->finalized, ONSAS_ERR_NOT_READY, "onsas_finalize_mesh has not been called");
    switch (c->dim) {
        case 1: run_cg_bs<1>(c, A); break;
        case 2: run_cg_bs<2>(c, A); break;
        default: run_cg_bs<3>(c, A); break;
    }
}

// Wires the peer-memory solver once every rank's window address is known (CUDA IPC mappings with one process per GPU,
// plain peer pointers when one process drives all devices): the push map -- for every owned dof the LL slots inside the
// neighbours' receive buffers that want its value -- the peers' scalar slots, and the interface bit of the dof mask.
void p2p_wire(onsas_ctx* c, const std::vector<unsigned char*>& win, const int64_t* remote_halo_off) {
    const int nn = (int)c->nbr_rank.size();
    std::vector<unsigned long long*> slots(c->n_ranks);
    for (int r = 0; r < c->n_ranks; ++r) slots[r] = win_slots(win[r]);
    const int bs = c->dim;
    const int64_t nd = c->n_own_dofs();
    std::vector<long long> pptr(nd + 1, 0);
    for (int k = 0; k < nn; ++k)
        for (int64_t j = c->send_ptr[k]; j < c->send_ptr[k + 1]; ++j)
            for (int q = 0; q < bs; ++q) pptr[(int64_t)c->h_send_nodes[j] * bs + q + 1]++;
    for (int64_t i = 0; i < nd; ++i) pptr[i + 1] += pptr[i];
    std::vector<unsigned long long*> pdst((size_t)pptr[nd]);
    std::vector<long long> fill(pptr.begin(), pptr.end() - 1);
    for (int k = 0; k < nn; ++k) {
        unsigned long long* zh = win_zh(win[c->nbr_rank[k]]);
        for (int64_t j = c->send_ptr[k]; j < c->send_ptr[k + 1]; ++j)
            for (int q = 0; q < bs; ++q) {
                const int64_t i = (int64_t)c->h_send_nodes[j] * bs + q;
                const int64_t remote_dof = (remote_halo_off[k] + (j - c->send_ptr[k])) * bs + q;
                pdst[fill[i]++] = zh + 2 * remote_dof;
            }
    }
    if (pdst.empty()) pdst.push_back(nullptr);
    cudaStream_t s = c->stream;
    c->d_push_ptr.upload(pptr, s);
    c->d_push_dst.upload(pdst, s);
    c->d_peer_slots.upload(slots, s);
    {
        std::vector<double*> pe(c->n_ranks);
        std::vector<unsigned long long*> pf(c->n_ranks), pw(c->n_ranks);
        for (int r = 0; r < c->n_ranks; ++r) {
            pe[r] = win_E2(win[r]);
            pf[r] = win_rowflags(win[r]);
            pw[r] = win_w2(win[r]);
        }
        c->d_peer_E2.upload(pe, s);
        c->d_peer_rowflags.upload(pf, s);
        c->d_peer_w2.upload(pw, s);
    }
    CUDA_CHECK(cudaStreamSynchronize(s));
    // mark the interface dofs (mask bit 1): only they look the push map up inside the solver
    c->h_iface.assign((size_t)nd, 0);
    for (int64_t i = 0; i < nd; ++i)
        if (pptr[i + 1] > pptr[i]) c->h_iface[i] = 1;
    upload_mask(c);
    c->p2p_ready = true;
}

void fetch_state(onsas_ctx* c) {
    CUDA_CHECK(cudaMemcpyAsync(c->h_st, c->st.p, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->h_flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

void check_deferred(onsas_ctx* c) {
    if (*c->h_flag != 0) {
        const int f = *c->h_flag;
        c->err_flag.zero(c->stream);
        *c->h_flag = 0;
        if (f == 2) throw OnsasError(ONSAS_ERR_COMM, "peer-memory CG: timed out waiting for another rank");
        if (f == 3) throw OnsasError(ONSAS_ERR_CUDA, "streamed CG: a shared-memory stage of K never arrived");
        if (f == 4) throw OnsasError(ONSAS_ERR_BREAKDOWN, "linear solve broke down: the CG residual is not finite (NaN in K, F or U)");
        throw OnsasError(ONSAS_ERR_NEGATIVE_VOLUME, "Element with negative volume, check connectivity.");
    }
}

void fill_info(onsas_ctx* c, onsas_step_info* info, float ms_a, float ms_s) {
    const CgState& s = *c->h_st;
    info->norm_dU = std::sqrt(s.dd);
    info->norm_U = std::sqrt(s.uu);
    info->norm_r = std::sqrt(s.rr0);
    info->norm_Fext = std::sqrt(s.ff);
    info->cg_iters = s.it;
    info->cg_residual = s.res;
    info->cg_tol = s.tol;
    info->ms_assemble = ms_a;
    info->ms_solve = ms_s;
}

void download(onsas_ctx* c, double* h, const double* d, size_t n) {
    CUDA_CHECK(cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int32_t onsas_version(void) { return 100; }

const char* onsas_last_error(onsas_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t onsas_create(int32_t device, onsas_ctx** out) {
    if (!out) return ONSAS_ERR_INVALID_ARG;
    *out = nullptr;
    onsas_ctx* c = nullptr;
    int32_t st = guard(nullptr, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw OnsasError(ONSAS_ERR_CUDA, std::string("no CUDA device available (libonsas_cuda has no CPU fallback): ") +
                                                 cudaGetErrorString(e));
        require(device >= 0 && device < ndev, ONSAS_ERR_INVALID_ARG, "device index out of range");
        CUDA_CHECK(cudaSetDevice(device));
        c = new onsas_ctx();
        c->device = device;
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        c->n_sm = prop.multiProcessorCount;
        require(prop.cooperativeLaunch != 0, ONSAS_ERR_UNSUPPORTED, "device lacks cooperative launch");
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        CUDA_CHECK(cudaMallocHost(&c->h_st, sizeof(CgState)));
        CUDA_CHECK(cudaMallocHost(&c->h_flag, sizeof(int)));
        std::memset(c->h_st, 0, sizeof(CgState));
        *c->h_flag = 0;
        for (auto& ev : c->ev) CUDA_CHECK(cudaEventCreate(&ev));
        c->err_flag.alloc(1);
        c->err_flag.zero(c->stream);
        c->st.alloc(1);
        c->st.zero(c->stream);
        c->partials.alloc((size_t)P_COUNT * c->part_stride);
        c->partials.zero(c->stream);
        c->red.alloc(8);
        c->red.zero(c->stream);
        c->prof.alloc(16 + 4096);
        c->prof.zero(c->stream);
    });
    if (st != ONSAS_OK) {
        delete c;
        return st;
    }
    *out = c;
    return ONSAS_OK;
}

int32_t onsas_destroy(onsas_ctx* c) {
    if (!c) return ONSAS_OK;
    if (c->grp) grp_destroy(c);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void* b : c->ipc_opened) cudaIpcCloseMemHandle(b);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (auto& ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->h_st) cudaFreeHost(c->h_st);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    drop_host_graph(c);
    if (c->hp.ev_join[0]) cudaEventDestroy(c->hp.ev_join[0]);
    if (c->hp.ev_join[1]) cudaEventDestroy(c->hp.ev_join[1]);
    if (c->hp.s_in) cudaStreamDestroy(c->hp.s_in);
    if (c->hp.s_out) cudaStreamDestroy(c->hp.s_out);
    if (c->hp.s_k2) cudaStreamDestroy(c->hp.s_k2);
    if (c->hp.ev_start) cudaEventDestroy(c->hp.ev_start);
    for (auto e : c->hp.ev_in) cudaEventDestroy(e);
    for (auto e : c->hp.ev_k) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return ONSAS_OK;
}

int32_t onsas_set_stream(onsas_ctx* c, void* s) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(!c->grp || c->grp->sub.size() == 1, ONSAS_ERR_UNSUPPORTED, "a multi-device context runs on its own streams (one per device)");
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        drop_host_graph(c);
        c->stream = s ? (cudaStream_t)s : c->own_stream;
        if (c->grp) {  // a renumbered single-device context: its device context does the work
            onsas_ctx* d = c->grp->sub[0];
            CUDA_CHECK(cudaStreamSynchronize(d->stream));
            d->stream = s ? (cudaStream_t)s : d->own_stream;
        }
    });
}

int32_t onsas_set_option(onsas_ctx* c, int32_t key, int64_t value) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        switch (key) {
            case ONSAS_OPT_CG_MODE: require(value >= 0 && value <= 2, ONSAS_ERR_INVALID_ARG, "cg mode must be 0, 1 or 2"); c->cg_mode = (int)value; break;
            case ONSAS_OPT_ASM_MINBLOCKS: require(value >= 1 && value <= 3, ONSAS_ERR_INVALID_ARG, "min blocks must be 1..3"); c->asm_minb = (int)value; break;
            case ONSAS_OPT_CG_CHECK_EVERY: require(value >= 1 && value <= 4096, ONSAS_ERR_INVALID_ARG, "check_every out of range"); c->check_every = (int)value; break;
            case ONSAS_OPT_FORCE_MG: c->force_mg = value != 0; break;
            case ONSAS_OPT_COARSE_FUSED: c->coarse_fused = value != 0; break;
            case ONSAS_OPT_COARSE_RBM: c->coarse_rbm = value != 0; c->co.built = false; c->co.fresh = false; break;
            case ONSAS_OPT_GJ_BLOCKED: c->gj_blocked = value != 0; c->co.fresh = false; break;
            case ONSAS_OPT_HOST_CHUNKS: require(value >= 1 && value <= 64, ONSAS_ERR_INVALID_ARG, "host chunks must be 1..64"); c->host_chunks = (int)value; c->hp.built = false; break;
            case ONSAS_OPT_HOST_STREAMS: require(value >= 1 && value <= 2, ONSAS_ERR_INVALID_ARG, "host streams must be 1 or 2"); c->host_streams = (int)value; break;
            case ONSAS_OPT_HOST_MID_WEIGHT: require(value >= 1 && value <= 64, ONSAS_ERR_INVALID_ARG, "weight must be 1..64"); c->host_mid_weight = (int)value; c->hp.built = false; break;
            case ONSAS_OPT_CG_PROFILE: c->cg_profile = value != 0; c->cg_grid = 0; break;
            case ONSAS_OPT_CG_BLOCKS_PER_SM: require(value >= 0 && value <= 32, ONSAS_ERR_INVALID_ARG, "blocks per SM out of range"); c->cg_bps = (int)value; c->cg_grid = 0; break;
            case ONSAS_OPT_COARSE_GLOBAL: c->coarse_global = value != 0; c->co.built = false; c->co.fresh = false; break;
            case ONSAS_OPT_HOST_GRAPH: c->host_graph = value != 0; c->hp.graph_failed = false; break;
            case ONSAS_OPT_TRUSS_MINBLOCKS: require(value >= 2 && value <= 4, ONSAS_ERR_INVALID_ARG, "truss min blocks must be 2..4"); c->truss_minb = (int)value; break;
            case ONSAS_OPT_CG_SINGLE_REDUCTION: require(value >= 0 && value <= 3, ONSAS_ERR_INVALID_ARG, "single-reduction mask must be 0..3"); c->cg_single_reduction = (int)value; break;
            case ONSAS_OPT_CG_L2_PREFETCH: require(value >= 0 && value <= 64, ONSAS_ERR_INVALID_ARG, "L2 prefetch distance must be 0..64 slices"); c->cg_l2_prefetch = (int)value; break;
            case ONSAS_OPT_REORDER: require(value >= 0 && value <= 2, ONSAS_ERR_INVALID_ARG, "reorder must be 0, 1 or 2"); require(!c->finalized, ONSAS_ERR_INVALID_ARG, "ONSAS_OPT_REORDER must be set before onsas_finalize_mesh"); c->reorder = (int)value; break;
            default: throw OnsasError(ONSAS_ERR_INVALID_ARG, "unknown option key");
        }
        c->opt_log.emplace_back(key, value);
        drop_host_graph(c);  // a captured pipeline bakes the launch configuration in
        if (c->grp && key != ONSAS_OPT_REORDER) grp_set_option(c, key, value);
    });
}

// ---------------------------------------------------------------- mesh upload
int32_t onsas_set_nodes(onsas_ctx* c, int64_t n_nodes, int64_t n_owned, int32_t dim, const double* xyz) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(dim >= 1 && dim <= 3, ONSAS_ERR_INVALID_ARG, "dim must be 1, 2 or 3");
        require(n_nodes >= 0 && n_owned >= 0 && n_owned <= n_nodes, ONSAS_ERR_INVALID_ARG, "bad node counts");
        require(n_nodes == 0 || xyz, ONSAS_ERR_INVALID_ARG, "xyz is NULL");
        require(n_nodes * dim < (int64_t)0x7fffffff, ONSAS_ERR_UNSUPPORTED, "more than 2^31 dofs per GPU are not supported");
        c->dim = dim;
        c->n_nodes = n_nodes;
        c->n_owned = n_owned;
        c->h_xyz.assign(xyz, xyz + n_nodes * dim);
        c->h_agg_ptr.clear();
        c->h_agg2.clear();
        c->h_cen2.clear();
        c->n_agg2_per_rank = 0;
        c->remote_halo_off.clear();
        c->have_nodes = true;
        c->finalized = false;
    });
}

int32_t onsas_set_materials(onsas_ctx* c, int32_t n, const int32_t* kind, const double* params) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(n > 0 && kind && params, ONSAS_ERR_INVALID_ARG, "bad material table");
        for (int i = 0; i < n; ++i) require(kind[i] >= 0 && kind[i] <= 2, ONSAS_ERR_INVALID_ARG, "unknown material kind");
        c->h_mat_kind.assign(kind, kind + n);
        c->h_mat_params.assign(params, params + 2 * n);
        if (c->finalized) {  // material swap on a finalized mesh (replace!(s, material), Structures.jl): the element kinds
            // are re-derived in place -- the tables, U, F_ext and the load patterns of the mesh stay as they are
            if (c->grp) return grp_materials_changed(c);
            derive_element_kinds(c);
            drop_host_graph(c);  // the element kind selects the kernel instantiation
            c->mat_kind.upload(c->h_mat_kind, c->stream);
            c->mat_params.upload(c->h_mat_params, c->stream);
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            c->co.fresh = false;
        }
    });
}

int32_t onsas_set_tets(onsas_ctx* c, int64_t n, const int32_t* conn, const int32_t* mat_id) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(n >= 0 && (n == 0 || conn), ONSAS_ERR_INVALID_ARG, "bad tetrahedron table");
        c->n_tets = n;
        c->h_tets.assign(conn, conn + 4 * n);
        c->tet_has_mat = mat_id != nullptr;
        if (mat_id) c->h_tet_mat.assign(mat_id, mat_id + n);
        else c->h_tet_mat.clear();
        c->finalized = false;
    });
}

int32_t onsas_set_trusses(onsas_ctx* c, int64_t n, const int32_t* conn, const int32_t* mat_id, const double* area,
                          int32_t strain_model) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(n >= 0 && (n == 0 || (conn && area)), ONSAS_ERR_INVALID_ARG, "bad truss table");
        require(strain_model == 0 || strain_model == 1, ONSAS_ERR_INVALID_ARG, "unknown strain model");
        c->n_trusses = n;
        c->h_trusses.assign(conn, conn + 2 * n);
        c->h_area.assign(area, area + n);
        c->truss_has_mat = mat_id != nullptr;
        if (mat_id) c->h_truss_mat.assign(mat_id, mat_id + n);
        else c->h_truss_mat.clear();
        c->strain_model = strain_model;
        c->finalized = false;
    });
}

int32_t onsas_set_free_dofs(onsas_ctx* c, int64_t n_free, const int64_t* free_dofs, int64_t n_free_global) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(c->have_nodes, ONSAS_ERR_NOT_READY, "onsas_set_nodes must be called first");
        require(n_free >= 0 && (n_free == 0 || free_dofs), ONSAS_ERR_INVALID_ARG, "bad free dof list");
        const int64_t nd = c->n_own_dofs(), nl = c->n_local_dofs();
        c->h_mask.assign((size_t)nl, 0);
        int64_t n_own_free = 0;
        for (int64_t k = 0; k < n_free; ++k) {  // dofs of halo nodes may be listed too: the solver then updates U on them like their owner does
            require(free_dofs[k] >= 0 && free_dofs[k] < nl, ONSAS_ERR_INVALID_ARG, "free dof out of range");
            c->h_mask[free_dofs[k]] = 1;
            n_own_free += free_dofs[k] < nd ? 1 : 0;
        }
        n_free = n_own_free;
        c->n_free = n_free;
        c->n_free_global = n_free_global > 0 ? n_free_global : n_free;
        c->have_free = true;
        if (c->finalized && c->grp) return grp_free_dofs_changed(c);
        if (c->finalized) upload_mask(c);
    });
}

int32_t onsas_finalize_mesh(onsas_ctx* c) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (!c->grp && c->reorder && c->n_ranks == 1 && c->n_owned == c->n_nodes) grp_wrap_single(c);
        if (c->grp) return grp_finalize(c);
        require(c->have_nodes, ONSAS_ERR_NOT_READY, "onsas_set_nodes has not been called");
        require(!c->h_mat_kind.empty(), ONSAS_ERR_NOT_READY, "onsas_set_materials has not been called");
        require(c->have_free, ONSAS_ERR_NOT_READY, "onsas_set_free_dofs has not been called");
        derive_element_kinds(c);
        std::string msg = build_mesh_tables(c->dim, c->n_nodes, c->n_owned, c->n_tets, c->h_tets.data(), c->n_trusses,
                                            c->h_trusses.data(), c->tab);
        if (!msg.empty()) throw OnsasError(ONSAS_ERR_INVALID_ARG, msg);
        // reference volume check (Tetrahedrons.jl:134-138), once: X never changes
        {
            bool bad = false;
            const double* X = c->h_xyz.data();
#pragma omp parallel for schedule(static) reduction(|| : bad)
            for (int64_t e = 0; e < c->n_tets; ++e) {
                const int32_t* nd = &c->h_tets[4 * e];
                double c0[3], c1[3], c2[3];
                for (int i = 0; i < 3; ++i) {
                    c0[i] = X[3 * (int64_t)nd[0] + i] - X[3 * (int64_t)nd[1] + i];
                    c1[i] = X[3 * (int64_t)nd[3] + i] - X[3 * (int64_t)nd[1] + i];
                    c2[i] = X[3 * (int64_t)nd[2] + i] - X[3 * (int64_t)nd[1] + i];
                }
                double det = c0[0] * (c1[1] * c2[2] - c1[2] * c2[1]) + c0[1] * (c1[2] * c2[0] - c1[0] * c2[2]) +
                             c0[2] * (c1[0] * c2[1] - c1[1] * c2[0]);
                if (!(det / 6.0 > 0.0)) bad = true;
            }
            if (bad) throw OnsasError(ONSAS_ERR_NEGATIVE_VOLUME, "Element with negative volume, check connectivity.");
        }
        const size_t max_smem = 227 * 1024;
        require(asm_smem_bytes(std::max(c->tab.fam[0].max_pairs_per_slice, 1), std::max(c->tab.max_width, 1), std::max(c->tab.fam[0].max_snodes, 1), TET_REC, 4, 3) <= max_smem &&
                    asm_smem_bytes(std::max(c->tab.fam[1].max_pairs_per_slice, 1), std::max(c->tab.max_width, 1), std::max(c->tab.fam[1].max_snodes, 1), truss_rec(c->dim), 2, c->dim) <= max_smem,
                ONSAS_ERR_UNSUPPORTED, "node valence too high: a slice of 8 nodes has more element pairs than fit in shared memory");

        // diagonal block position per row
        std::vector<int32_t> dslot((size_t)c->n_owned, 0);
        for (int64_t i = 0; i < c->n_owned; ++i) {
            const int64_t base = c->tab.slice_ptr[i / SLICE_ROWS];
            const int l = (int)(i % SLICE_ROWS);
            int s = 0;
            while (s < c->tab.row_nblk[i] && c->tab.col[(base + s) * SLICE_ROWS + l] != i) ++s;
            dslot[i] = s;
        }

        cudaStream_t s = c->stream;
        const size_t nl = (size_t)c->n_local_dofs(), no = (size_t)c->n_own_dofs();
        c->X.upload(c->h_xyz, s);
        c->mat_kind.upload(c->h_mat_kind, s);
        c->mat_params.upload(c->h_mat_params, s);
        c->tets.upload(c->h_tets, s);
        if (c->tet_has_mat) c->tet_mat.upload(c->h_tet_mat, s);
        c->trusses.upload(c->h_trusses, s);
        if (c->truss_has_mat) c->truss_mat.upload(c->h_truss_mat, s);
        c->area.upload(c->h_area, s);
        c->h_iface.clear();
        c->mask.upload(c->h_mask, s);
        c->slice_ptr.upload(c->tab.slice_ptr, s);
        c->col.upload(c->tab.col, s);
        c->diag_slot.upload(dslot, s);
        for (int f = 0; f < 2; ++f) {
            c->pair_code[f].upload(c->tab.fam[f].pair_code, s);
            c->snodes[f].upload(c->tab.fam[f].snodes, s);
            c->pair_lnodes[f].upload(c->tab.fam[f].pair_lnodes, s);
            c->hdr[f].upload(c->tab.fam[f].hdr, s);
            c->cptr[f].upload(c->tab.fam[f].cptr, s);
            c->ccode[f].upload(c->tab.fam[f].ccode, s);
        }
        c->val.alloc((size_t)c->tab.n_slots() * c->dim * c->dim);
        c->val.zero(s);
        c->U.alloc(nl); c->U.zero(s);
        c->Fext.alloc(nl); c->Fext.zero(s);
        c->Fint.alloc(nl); c->Fint.zero(s);
        c->p.alloc(nl);
        c->p.zero(s);
        c->p_pad.alloc((size_t)c->n_nodes * 4);  // cg_stream: one 32-byte sector per node
        c->p_pad.zero(s);
        if (c->n_ranks > 1 || c->force_mg) {
            // P2P window: fixed-size header (scalar slots, epochs) then the LL receive buffer of the halo dofs
            c->window.alloc(P2P_HDR_BYTES + P2P_COARSE_BYTES + std::max<size_t>(nl - no, 1) * 16);
            c->window.zero(s);
            c->p2p_ready = false;
        }
        c->rhs.alloc(nl); c->rhs.zero(s);
        c->x.alloc(nl); c->x.zero(s);  // owned + halo dofs: the peer-memory CG keeps the solution consistent on the halo
        c->r.alloc(no); c->r.zero(s);
        c->Ap.alloc(no); c->Ap.zero(s);
        c->s_vec.alloc(no); c->s_vec.zero(s);
        c->dinv.alloc(no); c->dinv.zero(s);
        c->tet_out.alloc((size_t)c->n_tets * 16); c->tet_out.zero(s);
        c->truss_out.alloc((size_t)c->n_trusses * 2); c->truss_out.zero(s);
        CUDA_CHECK(cudaStreamSynchronize(s));
        c->cg_grid = 0;
        c->st_plan.built = false;
        c->co.built = false;
        c->co.fresh = false;
        c->hp.built = false;
        c->n_patterns = 0;
        c->patterns.release();
        c->finalized = true;
    });
}

// ---------------------------------------------------------------- state vectors
int32_t onsas_set_U(onsas_ctx* c, const double* U) {
    if (!c || !U) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_set_vec(c, VEC_U, U);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        CUDA_CHECK(cudaMemcpyAsync(c->U.p, U, c->n_local_dofs() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the caller may free U right after return
    });
}
int32_t onsas_get_U(onsas_ctx* c, double* U) {
    if (!c || !U) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_vec(c, VEC_U, U);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        download(c, U, c->U.p, (size_t)c->n_local_dofs());
    });
}
int32_t onsas_set_Fext(onsas_ctx* c, const double* F) {
    if (!c || !F) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_set_vec(c, VEC_FEXT, F);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        CUDA_CHECK(cudaMemcpyAsync(c->Fext.p, F, c->n_local_dofs() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}
int32_t onsas_get_Fint(onsas_ctx* c, double* F) {
    if (!c || !F) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_vec(c, VEC_FINT, F);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        if (c->n_local_dofs() > c->n_own_dofs()) std::fill(F + c->n_own_dofs(), F + c->n_local_dofs(), 0.0);  // halo part
        // one stream synchronisation for the vector and the deferred-error flag
        CUDA_CHECK(cudaMemcpyAsync(c->h_flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        download(c, F, c->Fint.p, (size_t)c->n_own_dofs());
        check_deferred(c);
    });
}
int32_t onsas_get_dU(onsas_ctx* c, double* dU) {
    if (!c || !dU) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_vec(c, VEC_DU, dU);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        std::fill(dU, dU + c->n_local_dofs(), 0.0);
        download(c, dU, c->x.p, (size_t)c->n_own_dofs());
    });
}

// ---------------------------------------------------------------- external loads on the device
namespace {
// appends one zero-initialised pattern to the device array and returns its device pointer
double* append_pattern(onsas_ctx* c) {
    const size_t n = (size_t)c->n_local_dofs();
    DevBuf<double> grown;
    grown.alloc((size_t)(c->n_patterns + 1) * n);
    if (c->n_patterns > 0)
        CUDA_CHECK(cudaMemcpyAsync(grown.p, c->patterns.p, (size_t)c->n_patterns * n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_CHECK(cudaMemsetAsync(grown.p + (size_t)c->n_patterns * n, 0, n * sizeof(double), c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    std::swap(c->patterns.p, grown.p);
    std::swap(c->patterns.n, grown.n);
    c->n_patterns += 1;
    return c->patterns.p + (size_t)(c->n_patterns - 1) * n;
}
}  // namespace

int32_t onsas_add_face_load(onsas_ctx* c, int64_t n_faces, const int32_t* tri, int32_t kind, const double* values, int32_t* pattern_id) {
    if (!c || !pattern_id) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_add_face_load(c, n_faces, tri, kind, values, pattern_id);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        require(c->dim == 3, ONSAS_ERR_UNSUPPORTED, "face loads need a 3-D mesh");
        require(kind == 0 || kind == 1, ONSAS_ERR_INVALID_ARG, "unknown load kind");
        require(n_faces >= 0 && (n_faces == 0 || tri) && values, ONSAS_ERR_INVALID_ARG, "bad face list");
        const int64_t nn = c->n_nodes;
        std::vector<int64_t> ptr((size_t)nn + 1, 0);
        for (int64_t f = 0; f < n_faces; ++f)
            for (int k = 0; k < 3; ++k) {
                const int32_t nd = tri[3 * f + k];
                require(nd >= 0 && nd < nn, ONSAS_ERR_INVALID_ARG, "face node out of range");
                ptr[(size_t)nd + 1]++;
            }
        for (int64_t i = 0; i < nn; ++i) ptr[i + 1] += ptr[i];
        std::vector<int32_t> face((size_t)ptr[nn]);
        std::vector<int64_t> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t f = 0; f < n_faces; ++f)  // ascending face order inside every node's list
            for (int k = 0; k < 3; ++k) face[(size_t)fill[tri[3 * f + k]]++] = (int32_t)f;
        if (face.empty()) face.push_back(0);
        DevBuf<int64_t> d_ptr;
        DevBuf<int32_t> d_face, d_tri;
        d_ptr.upload(ptr, c->stream);
        d_face.upload(face, c->stream);
        std::vector<int32_t> htri(tri, tri + 3 * n_faces);
        if (htri.empty()) htri.push_back(0);
        d_tri.upload(htri, c->stream);
        double* F = append_pattern(c);
        if (nn > 0) {
            k_face_load<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>(c->X.p, d_tri.p, d_ptr.p, d_face.p, nn, kind, values[0],
                                                                              kind == 0 ? values[1] : 0.0, kind == 0 ? values[2] : 0.0, F);
            CUDA_CHECK(cudaGetLastError());
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        *pattern_id = c->n_patterns - 1;
    });
}

int32_t onsas_add_nodal_load(onsas_ctx* c, int64_t n, const int32_t* nodes, const double* values, int32_t* pattern_id) {
    if (!c || !pattern_id) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_add_nodal_load(c, n, nodes, values, pattern_id);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        require(n >= 0 && (n == 0 || nodes) && values, ONSAS_ERR_INVALID_ARG, "bad node list");
        std::vector<double> h((size_t)c->n_local_dofs(), 0.0);
        for (int64_t k = 0; k < n; ++k) {  // GlobalLoad on nodes (GlobalLoadBoundaryConditions.jl:34-48), duplicates summed in list order
            require(nodes[k] >= 0 && nodes[k] < c->n_nodes, ONSAS_ERR_INVALID_ARG, "load node out of range");
            for (int d = 0; d < c->dim; ++d) h[(size_t)nodes[k] * c->dim + d] += values[d];
        }
        double* F = append_pattern(c);
        if (!h.empty()) CUDA_CHECK(cudaMemcpyAsync(F, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        *pattern_id = c->n_patterns - 1;
    });
}

int32_t onsas_apply_loads(onsas_ctx* c, int32_t n_factors, const double* factors) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_apply_loads(c, n_factors, factors);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        require(n_factors == c->n_patterns && (n_factors == 0 || factors), ONSAS_ERR_INVALID_ARG,
                "one factor per load pattern is required");
        const int64_t n = c->n_local_dofs();
        if (n_factors == 0) {
            c->Fext.zero(c->stream);
            return;
        }
        c->factors.upload(factors, (size_t)n_factors, c->stream);
        k_combine_loads<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->patterns.p, c->factors.p, n_factors, n, c->Fext.p);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(c->stream));  // factors is caller-owned host memory
    });
}

int32_t onsas_clear_loads(onsas_ctx* c) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_clear_loads(c);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->n_patterns = 0;
        c->patterns.release();
    });
}

int32_t onsas_get_Fext(onsas_ctx* c, double* F) {
    if (!c || !F) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_vec(c, VEC_FEXT, F);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        download(c, F, c->Fext.p, (size_t)c->n_local_dofs());
    });
}

// ---------------------------------------------------------------- hot path
int32_t onsas_assemble(onsas_ctx* c) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {  // asynchronous; errors surface at the next synchronizing call
        if (c->grp) return grp_assemble(c);
        launch_assemble(c);
    });
}

int32_t onsas_assemble_host(onsas_ctx* c, const double* U, double* F_int) {
    if (!c || !U || !F_int) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_assemble_host(c, U, F_int);
        assemble_host(c, U, F_int);
    });
}

int32_t onsas_eval_elements(onsas_ctx* c, int32_t family, int64_t first, int64_t count, double* f, double* K, double* sig,
                            double* eps) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_eval_elements(c, family, first, count, f, K, sig, eps);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        require(family == 0 || family == 1, ONSAS_ERR_INVALID_ARG, "unknown element family");
        const int64_t ne = family == 0 ? c->n_tets : c->n_trusses;
        require(first >= 0 && count >= 0 && first + count <= ne, ONSAS_ERR_INVALID_ARG, "element range out of bounds");
        require(f && K && sig && eps, ONSAS_ERR_INVALID_ARG, "NULL output buffer");
        if (count == 0) return;
        const int nde = (family == 0 ? 4 : 2) * c->dim;
        DevBuf<double> df, dK, ds, de;
        df.alloc((size_t)count * nde);
        dK.alloc((size_t)count * nde * nde);
        ds.alloc((size_t)count * 9);
        de.alloc((size_t)count * 9);
        EvalArgs A{};
        A.first = first; A.count = count; A.dim = c->dim;
        A.X = c->X.p; A.U = c->U.p;
        A.conn = family == 0 ? c->tets.p : c->trusses.p;
        A.mat_id = family == 0 ? (c->tet_has_mat ? c->tet_mat.p : nullptr) : (c->truss_has_mat ? c->truss_mat.p : nullptr);
        A.mat_kind = c->mat_kind.p; A.mat_params = c->mat_params.p; A.area = c->area.p; A.strain_model = c->strain_model;
        A.f = df.p; A.K = dK.p; A.sig = ds.p; A.eps = de.p; A.err_flag = c->err_flag.p;
        const unsigned grid = (unsigned)((count + 127) / 128);
        if (family == 0) k_eval_tets<<<grid, 128, 0, c->stream>>>(A);
        else if (c->dim == 3) k_eval_trusses<3><<<grid, 128, 0, c->stream>>>(A);
        else if (c->dim == 2) k_eval_trusses<2><<<grid, 128, 0, c->stream>>>(A);
        else k_eval_trusses<1><<<grid, 128, 0, c->stream>>>(A);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(f, df.p, df.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(K, dK.p, dK.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(sig, ds.p, ds.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(eps, de.p, de.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(c->h_flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        check_deferred(c);
    });
}

// one Newton iteration in two halves, so that a multi-device context can enqueue every device's work before it waits for any
static void step_launch(onsas_ctx* c, bool assemble, int32_t precond, double reltol, double abstol, int64_t maxiter, int update_U) {
    require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
    require(precond >= 0 && precond <= 2, ONSAS_ERR_INVALID_ARG, "unknown preconditioner");
    CUDA_CHECK(cudaEventRecord(c->ev[0], c->stream));
    if (assemble) launch_assemble(c);
    CUDA_CHECK(cudaEventRecord(c->ev[1], c->stream));
    CgArgs A = make_cg_args(c, precond, reltol, abstol, maxiter, false, update_U);
    if (update_U == 2) A.rhs = c->Fext.p;  // the step! of a LinearStaticAnalysis: r = F_ext[free], U[free] = dU (LinearStaticAnalyses.jl:117-153)
    run_cg(c, A);
    CUDA_CHECK(cudaEventRecord(c->ev[2], c->stream));
}
static void step_finish(onsas_ctx* c, onsas_step_info* info) {
    fetch_state(c);
    float ms_a = 0, ms_s = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms_a, c->ev[0], c->ev[1]));
    CUDA_CHECK(cudaEventElapsedTime(&ms_s, c->ev[1], c->ev[2]));
    check_deferred(c);
    fill_info(c, info, ms_a, ms_s);
}
static void step_impl(onsas_ctx* c, bool assemble, int32_t precond, double reltol, double abstol, int64_t maxiter,
                      int update_U, onsas_step_info* info) {
    require(info != nullptr, ONSAS_ERR_INVALID_ARG, "info is NULL");
    if (c->grp) return grp_step(c, assemble, precond, reltol, abstol, maxiter, update_U, info);
    step_launch(c, assemble, precond, reltol, abstol, maxiter, update_U);
    step_finish(c, info);
}

int32_t onsas_newton_step(onsas_ctx* c, int32_t precond, double cg_reltol, double cg_abstol, int64_t cg_maxiter,
                          onsas_step_info* info) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] { step_impl(c, true, precond, cg_reltol, cg_abstol, cg_maxiter, 1, info); });
}

int32_t onsas_step(onsas_ctx* c, int32_t precond, double cg_reltol, double cg_abstol, int64_t cg_maxiter, int32_t update_U,
                   onsas_step_info* info) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        require(update_U >= 0 && update_U <= 2, ONSAS_ERR_INVALID_ARG, "update_U must be 0, 1 or 2");
        step_impl(c, false, precond, cg_reltol, cg_abstol, cg_maxiter, update_U, info);
    });
}

int32_t onsas_pcg(onsas_ctx* c, const double* b, double* x, int32_t precond, double reltol, double abstol, int64_t maxiter,
                  int64_t* iters, double* residual) {
    if (!c || !b || !x) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_pcg(c, b, x, precond, reltol, abstol, maxiter, iters, residual);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        require(precond >= 0 && precond <= 2, ONSAS_ERR_INVALID_ARG, "unknown preconditioner");
        CUDA_CHECK(cudaMemcpyAsync(c->rhs.p, b, c->n_local_dofs() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CgArgs A = make_cg_args(c, precond, reltol, abstol, maxiter, true, 0);
        run_cg(c, A);
        fetch_state(c);
        std::fill(x, x + c->n_local_dofs(), 0.0);
        download(c, x, c->x.p, (size_t)c->n_own_dofs());
        if (iters) *iters = c->h_st->it;
        if (residual) *residual = c->h_st->res;
    });
}

int32_t onsas_spmv(onsas_ctx* c, const double* x, double* y) {
    if (!c || !x || !y) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_spmv(c, x, y);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        CUDA_CHECK(cudaMemcpyAsync(c->p.p, x, c->n_local_dofs() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CgArgs A = make_cg_args(c, 0, 0, 0, 1, false, 0);
        const int Gr = spmv_grid(c);
        if (c->n_ranks > 1) halo_exchange(c, A.p, 0);
        switch (c->dim) {
            case 1: k_spmv_dot<1><<<Gr, CG_THREADS, 0, c->stream>>>(A, 0); break;
            case 2: k_spmv_dot<2><<<Gr, CG_THREADS, 0, c->stream>>>(A, 0); break;
            default: k_spmv_dot<3><<<Gr, CG_THREADS, 0, c->stream>>>(A, 0); break;
        }
        CUDA_CHECK(cudaGetLastError());
        std::fill(y, y + c->n_local_dofs(), 0.0);
        download(c, y, c->Ap.p, (size_t)c->n_own_dofs());
    });
}

/* Device-resident SpMV on whatever p currently holds (bench / profiling hook; asynchronous). */
int32_t onsas_spmv_resident(onsas_ctx* c) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_spmv_resident(c);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        CgArgs A = make_cg_args(c, 0, 0, 0, 1, false, 0);
        const int Gr = spmv_grid(c);
        switch (c->dim) {
            case 1: k_spmv_dot<1><<<Gr, CG_THREADS, 0, c->stream>>>(A, 0); break;
            case 2: k_spmv_dot<2><<<Gr, CG_THREADS, 0, c->stream>>>(A, 0); break;
            default: k_spmv_dot<3><<<Gr, CG_THREADS, 0, c->stream>>>(A, 0); break;
        }
        CUDA_CHECK(cudaGetLastError());
    });
}

/* Blocks until the context's stream is idle and reports deferred errors (negative volume). */
int32_t onsas_synchronize(onsas_ctx* c) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_synchronize(c);
        CUDA_CHECK(cudaMemcpyAsync(c->h_flag, c->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        check_deferred(c);
    });
}

// ---------------------------------------------------------------- results
int32_t onsas_get_csr_size(onsas_ctx* c, int64_t* n_rows, int64_t* nnz) {
    if (!c || !n_rows || !nnz) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_csr_size(c, n_rows, nnz);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        *n_rows = c->n_own_dofs();
        *nnz = c->tab.nnz_blocks * c->dim * c->dim;
    });
}

int32_t onsas_get_csr(onsas_ctx* c, int64_t* rowptr, int32_t* col, double* val) {
    if (!c || !rowptr || !col || !val) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_csr(c, rowptr, col, val);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        std::vector<int64_t> rp;
        std::vector<int32_t> ci;
        bsell_to_csr_pattern(c->tab, rp, ci);
        std::copy(rp.begin(), rp.end(), rowptr);
        std::copy(ci.begin(), ci.end(), col);
        std::vector<double> h(c->val.n);
        download(c, h.data(), c->val.p, c->val.n);
        bsell_to_csr_values(c->tab, h.data(), val);
    });
}

int32_t onsas_get_stress_strain(onsas_ctx* c, int32_t family, double* sig, double* eps) {
    if (!c || !sig || !eps) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_stress_strain(c, family, sig, eps);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        require(family == 0 || family == 1, ONSAS_ERR_INVALID_ARG, "unknown element family");
        if (family == 0) {
            // the staging buffer lives with the context (a fresh 128 MB vector per call cost more in page faults than the copy
            // itself: store! runs once per load step) and the records are unpacked by all host threads
            std::vector<double>& h = c->h_elem_stage;
            if (h.size() < c->tet_out.n) h.resize(c->tet_out.n);
            if (c->tet_out.n) download(c, h.data(), c->tet_out.p, c->tet_out.n);
            const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
#pragma omp parallel for schedule(static)
            for (int64_t e = 0; e < c->n_tets; ++e) {
                for (int k = 0; k < 9; ++k) sig[9 * e + k] = h[16 * e + k];
                for (int v = 0; v < 6; ++v) {
                    eps[9 * e + VI[v] + 3 * VJ[v]] = h[16 * e + 9 + v];
                    eps[9 * e + VJ[v] + 3 * VI[v]] = h[16 * e + 9 + v];
                }
            }
        } else {
            std::vector<double> h(c->truss_out.n);
            if (!h.empty()) download(c, h.data(), c->truss_out.p, h.size());
            for (int64_t e = 0; e < c->n_trusses; ++e) {
                for (int k = 0; k < 9; ++k) sig[9 * e + k] = eps[9 * e + k] = 0.0;
                sig[9 * e] = h[2 * e];
                eps[9 * e] = h[2 * e + 1];
            }
        }
    });
}

int32_t onsas_get_cg_profile(onsas_ctx* c, int64_t out[8]) {
    if (!c || !out) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return (void)onsas_get_cg_profile(c->grp->sub[0], out);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::vector<long long> h(16 + 4096);
        CUDA_CHECK(cudaMemcpy(h.data(), c->prof.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 7; ++k) out[k] = h[k];
        if (h[7] > 0) {  // multi-GPU variant: [3] is split into the grid sync ([7]) and the reduction proper
            out[3] = h[3];
            out[6] += 0;
        }
        // [7]: slowest CTA's accumulated SpMV cycles (multi-GPU profiling variant), 0 otherwise
        long long mx = 0;
        for (size_t k = 16; k < 16 + 2048; ++k) mx = std::max(mx, h[k]);
        out[7] = mx;
        if (getenv("ONSAS_PROF_VERBOSE"))
            fprintf(stderr, "[onsas prof] two-level coarse apply, cycles of block 0 (whole solve): barrier %lld, w = Z^T r %lld, barrier %lld, y = E^-1 w %lld\n", h[8], h[9], h[10], h[11]);
        if (const char* f = getenv("ONSAS_PROF_DUMP")) {  // diagnostics: per-CTA SpMV cycles of the last profiled solve
            if (FILE* fp = fopen(f, "w")) {
                for (size_t k = 16; k < h.size(); ++k) fprintf(fp, "%lld\n", h[k]);
                fclose(fp);
            }
        }
        c->prof.zero(c->stream);
    });
}

int32_t onsas_get_table_stats(onsas_ctx* c, int64_t out[8]) {
    if (!c || !out) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return grp_get_table_stats(c, out);
        require(c->finalized, ONSAS_ERR_NOT_READY, "mesh not finalized");
        out[0] = c->tab.n_slices;
        out[1] = c->tab.n_slots();
        out[2] = c->tab.nnz_blocks;
        out[3] = (int64_t)c->tab.fam[0].pair_code.size();
        out[4] = (int64_t)c->tab.fam[1].pair_code.size();
        out[5] = std::max(c->tab.fam[0].max_pairs_per_slice, c->tab.fam[1].max_pairs_per_slice);
        out[6] = (int64_t)(c->val.n * sizeof(double));
        if (c->cg_grid == 0) {
            switch (c->dim) {
                case 1: c->cg_grid = persistent_grid<1>(c); break;
                case 2: c->cg_grid = persistent_grid<2>(c); break;
                default: c->cg_grid = persistent_grid<3>(c); break;
            }
        }
        out[7] = c->cg_grid;
    });
}

// ---------------------------------------------------------------- multi-GPU
int32_t onsas_comm_unique_id(void* id128) {
    if (!id128) return ONSAS_ERR_INVALID_ARG;
    std::string err;
    if (!g_nccl.load(err)) {
        g_create_error = err;
        return ONSAS_ERR_COMM;
    }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ONSAS_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return ONSAS_OK;
}

int32_t onsas_comm_init(onsas_ctx* c, int32_t n_ranks, int32_t rank, const void* id128) {
    if (!c || !id128) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return throw OnsasError(ONSAS_ERR_INVALID_ARG, "a multi-device context manages its own ranks");
        require(n_ranks >= 1 && rank >= 0 && rank < n_ranks, ONSAS_ERR_INVALID_ARG, "bad rank / n_ranks");
        std::string err;
        if (!g_nccl.load(err)) throw OnsasError(ONSAS_ERR_COMM, err);
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        NCCL_CHECK(g_nccl.CommInitRank(&c->comm, n_ranks, id, rank));
        c->n_ranks = n_ranks;
        c->rank = rank;
    });
}

int32_t onsas_set_halo(onsas_ctx* c, int32_t n_nbr, const int32_t* nbr_rank, const int64_t* send_ptr,
                       const int32_t* send_nodes, const int64_t* recv_ptr) {
    if (!c) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return throw OnsasError(ONSAS_ERR_INVALID_ARG, "a multi-device context manages its own halo plan");
        require(n_nbr >= 0, ONSAS_ERR_INVALID_ARG, "bad neighbour count");
        require(n_nbr == 0 || (nbr_rank && send_ptr && recv_ptr), ONSAS_ERR_INVALID_ARG, "NULL halo arrays");
        c->nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
        c->send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
        c->recv_ptr.assign(recv_ptr, recv_ptr + n_nbr + 1);
        const int64_t ns = n_nbr ? send_ptr[n_nbr] : 0;
        require(ns == 0 || send_nodes, ONSAS_ERR_INVALID_ARG, "NULL send list");
        for (int64_t k = 0; k < ns; ++k)
            require(send_nodes[k] >= 0 && send_nodes[k] < c->n_owned, ONSAS_ERR_INVALID_ARG, "send node is not owned");
        require((n_nbr ? recv_ptr[n_nbr] : 0) == c->n_nodes - c->n_owned, ONSAS_ERR_INVALID_ARG,
                "receive ranges do not cover the halo nodes");
        c->h_send_nodes.assign(send_nodes, send_nodes + ns);
        c->send_nodes.upload(send_nodes, (size_t)ns, c->stream);
        c->sendbuf.alloc((size_t)std::max<int64_t>(ns, 1) * c->dim);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

int32_t onsas_p2p_export(onsas_ctx* c, void* handle64, int64_t* offset) {
    if (!c || !handle64 || !offset) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return throw OnsasError(ONSAS_ERR_INVALID_ARG, "a multi-device context wires its own peer windows");
        require(c->finalized && c->window.p, ONSAS_ERR_NOT_READY,
                "onsas_comm_init and onsas_finalize_mesh must precede onsas_p2p_export");
        cudaIpcMemHandle_t h;
        CUDA_CHECK(cudaIpcGetMemHandle(&h, c->window.p));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        std::memcpy(handle64, &h, 64);
        // the handle names the whole underlying allocation: report where the window starts inside it
        *offset = 0;
        void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (drv) {
            using Fn = int (*)(unsigned long long*, size_t*, unsigned long long);
            Fn f = (Fn)dlsym(drv, "cuMemGetAddressRange_v2");
            unsigned long long base = 0;
            size_t size = 0;
            if (f && f(&base, &size, (unsigned long long)(uintptr_t)c->window.p) == 0)
                *offset = (int64_t)((unsigned long long)(uintptr_t)c->window.p - base);
        }
    });
}

int32_t onsas_p2p_import(onsas_ctx* c, const void* handles, const int64_t* offsets, const int64_t* remote_halo_off) {
    if (!c || !handles || !offsets) return ONSAS_ERR_INVALID_ARG;
    return guard(c, [&] {
        if (c->grp) return throw OnsasError(ONSAS_ERR_INVALID_ARG, "a multi-device context wires its own peer windows");
        require(c->finalized && c->window.p, ONSAS_ERR_NOT_READY, "window not allocated");
        require(c->n_ranks <= P2P_MAXR, ONSAS_ERR_UNSUPPORTED, "peer-memory CG supports at most 16 ranks");
        const int nn = (int)c->nbr_rank.size();
        // a context loaded from a partition (onsas_part_load) knows where its values start inside every neighbour's halo
        const int64_t* rho = remote_halo_off ? remote_halo_off : (c->remote_halo_off.size() == (size_t)nn ? c->remote_halo_off.data() : nullptr);
        require(nn == 0 || rho, ONSAS_ERR_INVALID_ARG, "NULL remote halo offsets");
        std::vector<unsigned char*> win(c->n_ranks, nullptr);
        for (int r = 0; r < c->n_ranks; ++r) {
            if (r == c->rank) {
                win[r] = c->window.p;
                continue;
            }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const unsigned char*)handles + 64 * r, 64);
            void* base = nullptr;
            CUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            c->ipc_opened.push_back(base);
            win[r] = (unsigned char*)base + offsets[r];
        }
        p2p_wire(c, win, rho);
    });
}

}  // extern "C"

#include "group.inc"
