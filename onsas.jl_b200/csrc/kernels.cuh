// kernels.cuh -- sm_100a kernels of the Newton-Raphson hot path.
//
//   k_assemble<FAMILY,KIND,DIM,ACCUM>  fused element evaluation + deterministic row-owner assembly
//                                      (replaces StaticAnalyses.jl:99-132 + Assemblers.jl:46-88)
//   k_eval_tets / k_eval_trusses       un-assembled f_e, K_e, sigma, eps (Entities.jl:174-218 contract)
//   cg_persistent<BS>                  whole Jacobi-PCG solve + residual + update + norms in ONE
//                                      cooperative launch (NonLinearStaticAnalyses.jl:107-148 and the
//                                      IterativeSolvers.jl cg! it calls); with N GPUs the same kernel pushes
//                                      halo values and partial sums into the peers' memory over NVLink
//   k_cg_* / k_spmv_dot<BS>            the same phases as separate launches (multi-GPU path, where the
//                                      halo exchange and all-reduce sit between them)
//
// All FP64, no tensor cores: nothing here is a dense contraction.  HBM-bound kernels are laid out
// for full-sector coalesced access (BSELL slices of 8 rows = 64-byte segments, 128-byte element
// output records); every reduction has a fixed order, so results are bitwise run-to-run reproducible.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>
#include <type_traits>

#include "element_math.cuh"
#include "tables.hpp"

namespace onsas {
namespace cg = cooperative_groups;

template <int V>
struct IC {
    static constexpr int value = V;
};

constexpr int C = SLICE_ROWS;

// ------------------------------------------------------------------------------------------------
// assembly
// ------------------------------------------------------------------------------------------------
struct AsmArgs {
    const double* X;  // dim per node
    const double* U;  // dim per node
    const SliceHdr* hdr;         // one 48-byte header per slice
    const int32_t* snodes;       // per slice: the distinct nodes its elements touch (ascending)
    const uint16_t* pair_lnodes; // npe indices into the slice's node list per pair (bit 15 of the first: writes the element record)
    const int32_t* pair_code;    // e*npe + a per pair
    const int32_t* mat_id;       // may be null (all elements use material 0)
    const int32_t* mat_kind;
    const double* mat_params;  // 2 per material
    const double* area;        // trusses
    int strain_model;          // trusses
    const uint32_t* cptr;
    const uint16_t* ccode;
    double* val;
    double* F_int;
    double* elem_out;  // tets: 16 doubles per element; trusses: 2 per element
    int* err_flag;     // set to 1 when an element has non-positive volume
    int max_pairs;     // sizes of the shared-memory regions
    int max_width;
    int max_snodes;
    int64_t n_rows_guard;  // number of owned rows (the last slice may be partial)
    int slice0;            // first slice of this launch (CTA b works on slice slice0 + b): onsas_assemble_host launches ranges
};

// shared-memory record of one staged node: X then U, padded to an odd number of doubles (16 consecutive records sit on 16
// different bank pairs, so a half-warp of 8-byte reads to distinct nodes of a run is conflict-free)
__host__ __device__ constexpr int snode_rec(int dim) { return 2 * dim + 1; }

// material of a pair's element, fetched together with the node data (before the staging barrier, not after it)
struct PairMat {
    double p0, p1;  // the two material parameters
    double area;    // trusses
    int kind;
};
template <int FAMILY>
__device__ __forceinline__ PairMat pair_material(const AsmArgs& A, int32_t code) {
    const int64_t e = FAMILY == 0 ? code >> 2 : code >> 1;
    const int m = A.mat_id ? __ldg(A.mat_id + e) : 0;
    PairMat M;
    M.p0 = __ldg(A.mat_params + 2 * m);
    M.p1 = __ldg(A.mat_params + 2 * m + 1);
    M.kind = __ldg(A.mat_kind + m);
    M.area = FAMILY == 1 ? __ldg(A.area + e) : 0.0;
    return M;
}

template <int KIND>
__device__ __forceinline__ void tet_pair(const AsmArgs& A, const double* sn, const uint2 ln, int32_t code, const PairMat& M, double* rec) {
    const int64_t e = code >> 2;
    const int a = code & 3;
    const unsigned li[4] = {ln.x & 0x7fffu, ln.x >> 16, ln.y & 0xffffu, ln.y >> 16};
    const bool writer = (ln.x & 0x8000u) != 0;
    double X[4][3], U[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double* np = sn + li[k] * snode_rec(3);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            X[k][c] = np[c];
            U[k][c] = np[3 + c];
        }
    }
    const double p0 = M.p0, p1 = M.p1;
    int kind = KIND;
    if (KIND == MAT_MIXED) kind = M.kind;

    double vol;
    auto run = [&](auto tag) {
        constexpr int K = decltype(tag)::value;
        TetCommon c;
        tet_common<K>(X, U, p0, p1, c);
        vol = c.vol;
        if (writer) {  // the pair of the element's first owned node also writes its stress / strain record
            double out[16];
            tet_stress_out<K>(c, p0, p1, out);
            // four 32-byte stores per 128-byte record (sm_100: 256-bit global stores): the writer lanes of a warp sit on
            // different lines, so every store instruction costs one L1 wavefront per writer -- half as many as with 16-byte stores
            double* o = A.elem_out + 16 * e;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * k), "d"(out[4 * k]), "d"(out[4 * k + 1]), "d"(out[4 * k + 2]),
                             "d"(out[4 * k + 3])
                             : "memory");
        }
        tet_row<K>(c, U, a, rec, rec + 36);
    };
    if (kind == MAT_SVK)
        run(IC<MAT_SVK>());
    else if (kind == MAT_NEOHOOKEAN)
        run(IC<MAT_NEOHOOKEAN>());
    else
        run(IC<MAT_ISOLINEAR>());
    if (!(vol > 0.0)) *A.err_flag = 1;  // Tetrahedrons.jl:134-138
}

template <int DIM>
__device__ __forceinline__ void truss_pair(const AsmArgs& A, const double* sn, const uint32_t ln, int32_t code, const PairMat& M, double* rec) {
    const int64_t e = code >> 1;
    const int a = code & 1;
    const double* n0 = sn + (ln & 0x7fffu) * snode_rec(DIM);
    const double* n1 = sn + (ln >> 16) * snode_rec(DIM);
    const bool writer = (ln & 0x8000u) != 0;
    double X[2][3], U[2][3];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        X[0][c] = n0[c];
        X[1][c] = n1[c];
        U[0][c] = n0[DIM + c];
        U[1][c] = n1[DIM + c];
    }
    const double Emod = truss_modulus(M.kind, M.p0, M.p1);
    double blk[2][9], f[3], se[2];
    truss_row<DIM>(A.strain_model, X, U, Emod, M.area, a, blk, f, se);
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int k = 0; k < DIM * DIM; ++k) rec[b * DIM * DIM + k] = blk[b][k];
#pragma unroll
    for (int r = 0; r < DIM; ++r) rec[2 * DIM * DIM + r] = f[r];
    if (writer) *reinterpret_cast<double2*>(A.elem_out + 2 * e) = make_double2(se[0], se[1]);  // one 16-byte store per record
}

// shared memory of one assembly CTA:
// [stage: max_pairs*REC (+ skew) doubles][ASM_ZEROS zeros][nodes: max_snodes * (2 dim + 1) doubles][scode: max_pairs*NPE u16][scp: max_width*C+1 u16]
constexpr int ASM_ZEROS = 4;  // >= DIM doubles, keeps the node records 32-byte aligned
__host__ __device__ constexpr size_t asm_smem_bytes(int max_pairs, int max_width, int max_snodes, int rec, int npe, int dim) {
    return ((size_t)max_pairs * rec + ROW_SKEW * SLICE_ROWS + ASM_ZEROS + (size_t)max_snodes * snode_rec(dim)) * 8 +
           (((size_t)max_pairs * npe * 2 + ((size_t)max_width * SLICE_ROWS + 1) * 2 + 15) / 16) * 16;
}

// One CTA per BSELL slice (8 block rows).
//  Staging: the slice's node list (the distinct nodes its elements touch, ascending ids -> runs of consecutive nodes) is read
//           once and X, U of those nodes land in shared memory with coalesced loads; the pairs address them by 16-bit index.
//           Four dependent load levels: slice header -> {node list, pair records, contribution codes, slot ranges} ->
//           {X, U of the listed nodes} -> barrier.
//  Phase A: one thread per (row, element) pair: evaluates block-row a of K_e and f_a from the staged nodes straight into
//           its shared-memory record and stages the slice's contribution lists in shared memory on the way.
//  Phase B: one thread per (block slot, block row r): sums the DIM entries of that row of the block over the
//           slot's contributions in ascending element order (register accumulators, indices from shared memory)
//           and writes K exactly once in 64-byte segments; the last C*DIM items do the same for F_int.
//  ACCUM adds onto what a previous family already wrote.
template <int FAMILY, int KIND, int DIM, bool ACCUM>
__device__ __forceinline__ void assemble_body(const AsmArgs& A) {
    extern __shared__ double stage[];
    constexpr int BB = DIM * DIM;
    constexpr int NPE = FAMILY == 0 ? 4 : 2;
    constexpr int REC = FAMILY == 0 ? TET_REC : truss_rec(DIM);
    constexpr int FOFF = NPE * BB;
    constexpr int NS = snode_rec(DIM);
    const int zero_off = A.max_pairs * REC + ROW_SKEW * C;  // ASM_ZEROS doubles of zeros: what the batched sums of phase B read past a list's end
    double* const snd = stage + zero_off + ASM_ZEROS;
    if (threadIdx.x < ASM_ZEROS) stage[zero_off + threadIdx.x] = 0.0;
    uint16_t* scode = reinterpret_cast<uint16_t*>(snd + (size_t)A.max_snodes * NS);
    uint16_t* scp = scode + (size_t)A.max_pairs * NPE;
    const int tid = threadIdx.x, nth = blockDim.x;

    // ---- level 1: the slice header (same address for every thread: one sector, broadcast)
    const int slice = A.slice0 + (int)blockIdx.x;
    const int4* hp = reinterpret_cast<const int4*>(A.hdr + slice);
    const int4 h0 = __ldg(hp), h1 = __ldg(hp + 1), h2 = __ldg(hp + 2);
    const int64_t p0 = (int64_t)(uint32_t)h0.x | ((int64_t)h0.y << 32);
    const int64_t base = (int64_t)(uint32_t)h0.z | ((int64_t)h0.w << 32);
    const int np = h1.x, width = h1.y;
    const int nsn = (int)(((uint32_t)h2.z >> 16) & 0xffffu);
    const uint32_t sn0 = (uint32_t)h2.w;
    const uint32_t cbase = (uint32_t)(p0 * NPE);
    const int nscp = width * C + 1;
    // The header is CTA-uniform: phase B re-reads it from shared memory (one STS here) instead of every thread carrying
    // it in registers -- or, at 96 registers, in per-thread local memory -- across the element arithmetic.
    __shared__ int4 s_hdr[3];
    if (tid == 0) {
        s_hdr[0] = h0;
        s_hdr[1] = h1;
        s_hdr[2] = h2;
    }
    auto row_off_of = [](const int4& a1, const int4& a2, int l) -> int {
        const uint32_t w = l < 2 ? (uint32_t)a1.z : l < 4 ? (uint32_t)a1.w : l < 6 ? (uint32_t)a2.x : l < 8 ? (uint32_t)a2.y : (uint32_t)a2.z;
        return (int)((w >> (16 * (l & 1))) & 0xffffu);
    };
    auto row_of_pair = [&](const int4& a1, const int4& a2, int t) -> int {  // row of pair t inside the slice
        int l = 0;
#pragma unroll
        for (int k = 1; k < C; ++k) l += (t >= row_off_of(a1, a2, k)) ? 1 : 0;
        return l;
    };

    // ---- level 2 (independent loads, issued together): node list, pair records, contribution codes, slot ranges
    using NodeVec = typename std::conditional<FAMILY == 0, uint2, uint32_t>::type;  // NPE u16 local node indices
    using CodeVec = typename std::conditional<FAMILY == 0, uint2, uint32_t>::type;  // NPE u16 codes of "pair t's chunk"
    const NodeVec* pn = reinterpret_cast<const NodeVec*>(A.pair_lnodes) + p0;
    const CodeVec* cc = reinterpret_cast<const CodeVec*>(A.ccode + cbase);
    int t = tid;
    NodeVec nodes = NodeVec();
    CodeVec chunk = CodeVec();
    int32_t code = 0;
    // staging runs one thread per (listed node, component): consecutive threads read consecutive doubles inside every run of
    // consecutive node ids (one thread per node read 24-byte strides: three instructions per vector over the same lines)
    // tets: staging runs one thread per (listed node, component); trusses: one thread per listed node (their lists are longer
    // than the CTA has threads once split by component: 7.6 against 8.5 G bars/s on the 10 M-bar lattice, profiles/r80)
    constexpr int SD = FAMILY == 0 ? DIM : 1;  // list entries are split into SD parts
    const int nsd = nsn * SD;
    int64_t gnode = -1;
    if (tid < nsd) gnode = __ldg(A.snodes + sn0 + tid / SD);
    if (t < np) {
        nodes = __ldg(pn + t);
        code = __ldg(A.pair_code + p0 + t);
        chunk = __ldg(cc + t);
    }
    uint32_t cp0 = 0;
    if (tid < nscp) cp0 = __ldg(A.cptr + base * C + tid);
    // ---- level 3: X, U of the listed nodes -> shared memory (consecutive threads hold consecutive list entries)
    {
        if constexpr (FAMILY == 0) {
            if (gnode >= 0) {
                const int c = tid % DIM;
                const double xv = __ldg(A.X + gnode * DIM + c), uv = __ldg(A.U + gnode * DIM + c);
                snd[(tid / DIM) * NS + c] = xv;
                snd[(tid / DIM) * NS + DIM + c] = uv;
            }
            for (int i = tid + nth; i < nsd; i += nth) {  // the rest of the list (270 dofs on the structured tet mesh, 192 threads)
                const int64_t g = __ldg(A.snodes + sn0 + i / DIM);
                const int c = i % DIM;
                snd[(i / DIM) * NS + c] = __ldg(A.X + g * DIM + c);
                snd[(i / DIM) * NS + DIM + c] = __ldg(A.U + g * DIM + c);
            }
        } else {
            double xv[DIM], uv[DIM];
            if (gnode >= 0) {
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    xv[c] = __ldg(A.X + gnode * DIM + c);
                    uv[c] = __ldg(A.U + gnode * DIM + c);
                }
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    snd[tid * NS + c] = xv[c];
                    snd[tid * NS + DIM + c] = uv[c];
                }
            }
            for (int i = tid + nth; i < nsn; i += nth) {  // lists longer than the CTA (high-valence meshes)
                const int64_t g = __ldg(A.snodes + sn0 + i);
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    snd[i * NS + c] = __ldg(A.X + g * DIM + c);
                    snd[i * NS + DIM + c] = __ldg(A.U + g * DIM + c);
                }
            }
        }
    }
    PairMat mat = PairMat();
    if (t < np) mat = pair_material<FAMILY>(A, code);
    if (tid < nscp) scp[tid] = (uint16_t)(cp0 - cbase);
    for (int i = tid + nth; i < nscp; i += nth) scp[i] = (uint16_t)(__ldg(A.cptr + base * C + i) - cbase);
    __syncthreads();

    // ---- phase A
    int l = row_of_pair(h1, h2, t);  // the pair's record is skewed by its row inside the slice (bank spreading)
    const int np_a = np;
    while (t < np_a) {
        reinterpret_cast<CodeVec*>(scode)[t] = chunk;
        if constexpr (FAMILY == 0)
            tet_pair<KIND>(A, snd, nodes, code, mat, stage + (size_t)t * REC + row_skew(FAMILY, l));
        else
            truss_pair<DIM>(A, snd, nodes, code, mat, stage + (size_t)t * REC + row_skew(FAMILY, l));
        t += nth;
        if (t < np_a) {  // slices with more pairs than threads (high-valence meshes): header again from L1, not from registers
            const int4* hq = reinterpret_cast<const int4*>(A.hdr + slice);
            const int64_t pq = (int64_t)(uint32_t)__ldg(hq).x | ((int64_t)__ldg(hq).y << 32);
            nodes = __ldg(reinterpret_cast<const NodeVec*>(A.pair_lnodes) + pq + t);
            code = __ldg(A.pair_code + pq + t);
            chunk = __ldg(reinterpret_cast<const CodeVec*>(A.ccode + (uint32_t)(pq * NPE)) + t);
            mat = pair_material<FAMILY>(A, code);
            l = row_of_pair(__ldg(hq + 1), __ldg(hq + 2), t);
        }
    }
    __syncthreads();

    // ---- phase B (header from shared memory)
    const int4 k0 = s_hdr[0], k1 = s_hdr[1], k2 = s_hdr[2];
    const int64_t base_b = (int64_t)(uint32_t)k0.z | ((int64_t)k0.w << 32);
    const int width_b = k1.y;
    auto row_off = [&](int r) -> int { return row_off_of(k1, k2, r); };
    const int nK = width_b * DIM * C;
    const int nF = C * DIM;
    double* const vout = A.val + base_b * BB * C;
    for (int w = tid; w < nK + nF; w += nth) {
        if (w < nK) {
            const int lane = w % C;
            const int r = (w / C) % DIM;
            const int s = w / (C * DIM);
            const int slot = s * C + lane;
            const int q0 = scp[slot], q1 = scp[slot + 1];
            double acc[DIM];
#pragma unroll
            for (int j = 0; j < DIM; ++j) acc[j] = 0.0;
            // contributions in batches of PB: the PB code loads, then the PB * DIM value loads are independent of each other
            // (two dependent shared-memory round trips per batch instead of per contribution); the ADDS keep the ascending
            // element order.  Batch slots past the end read the CTA's block of zeros.
            constexpr int PB = FAMILY == 0 ? 4 : 2;  // tets: 6.4 contributions per slot on the structured mesh; trusses: 1.5
            for (int q = q0; q < q1; q += PB) {
                int cd[PB];
#pragma unroll
                for (int k = 0; k < PB; ++k) cd[k] = q + k < q1 ? (int)scode[q + k] + r * DIM : zero_off;
                double v[PB][DIM];
#pragma unroll
                for (int k = 0; k < PB; ++k)
#pragma unroll
                    for (int j = 0; j < DIM; ++j) v[k][j] = stage[cd[k] + j];
#pragma unroll
                for (int k = 0; k < PB; ++k)
#pragma unroll
                    for (int j = 0; j < DIM; ++j) acc[j] += v[k][j];
            }
            double* dst = vout + ((size_t)s * BB + r * DIM) * C + lane;
#pragma unroll
            for (int j = 0; j < DIM; ++j) {
                if (ACCUM) acc[j] += dst[j * C];
                dst[j * C] = acc[j];
            }
        } else {
            const int j = w - nK;
            const int lane = j / DIM, r = j % DIM;
            const int t0 = row_off(lane), t1 = row_off(lane + 1);
            const int64_t row = (int64_t)slice * C + lane;
            if (t1 > t0 || !ACCUM) {
                double acc = 0.0;
                constexpr int FB = 8;  // the same batching for the row's force entries (24 pairs per row on the structured mesh)
                for (int tt = t0; tt < t1; tt += FB) {
                    double v[FB];
#pragma unroll
                    for (int k = 0; k < FB; ++k) v[k] = stage[tt + k < t1 ? (tt + k) * REC + row_skew(FAMILY, lane) + FOFF + r : zero_off];
#pragma unroll
                    for (int k = 0; k < FB; ++k) acc += v[k];
                }
                if (row < A.n_rows_guard) {
                    if (ACCUM) acc += A.F_int[row * DIM + r];
                    A.F_int[row * DIM + r] = acc;
                }
            }
        }
    }
}

// Register-budget variants of the same body (ONSAS_OPT_ASM_MINBLOCKS): launch bounds (256,1) / (256,2), or an
// explicit register cap (96 = 3 resident CTAs of 192 threads (the structured-mesh slice size).
template <int FAMILY, int KIND, int DIM, bool ACCUM, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_assemble(AsmArgs A) {
    assemble_body<FAMILY, KIND, DIM, ACCUM>(A);
}
template <int FAMILY, int KIND, int DIM, bool ACCUM, int REGS>
__global__ void __maxnreg__(REGS) k_assemble_reg(AsmArgs A) {
    assemble_body<FAMILY, KIND, DIM, ACCUM>(A);
}

// ------------------------------------------------------------------------------------------------
// un-assembled element evaluation (parity API)
// ------------------------------------------------------------------------------------------------
struct EvalArgs {
    int64_t first, count;
    int dim;
    const double* X;
    const double* U;
    const int32_t* conn;
    const int32_t* mat_id;
    const int32_t* mat_kind;
    const double* mat_params;
    const double* area;
    int strain_model;
    double* f;    // ndof_e per element
    double* K;    // ndof_e^2 per element, column-major
    double* sig;  // 9 per element, column-major
    double* eps;  // 9 per element
    int* err_flag;
};

__global__ void k_eval_tets(EvalArgs A) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= A.count) return;
    const int64_t e = A.first + i;
    double X[4][3], U[4][3];
    for (int k = 0; k < 4; ++k) {
        const int64_t nd = A.conn[4 * e + k];
        for (int c = 0; c < 3; ++c) {
            X[k][c] = A.X[3 * nd + c];
            U[k][c] = A.U[3 * nd + c];
        }
    }
    const int m = A.mat_id ? A.mat_id[e] : 0;
    const double p0 = A.mat_params[2 * m], p1 = A.mat_params[2 * m + 1];
    const int kind = A.mat_kind[m];
    double* K = A.K + 144 * i;
    double* f = A.f + 12 * i;
    double out[16];
    double vol = 0;
    auto run = [&](auto tag) {
        constexpr int KD = decltype(tag)::value;
        TetCommon c;
        tet_common<KD>(X, U, p0, p1, c);
        vol = c.vol;
        for (int a = 0; a < 4; ++a) {
            double blk[36], fa[3];
            tet_row<KD>(c, U, a, blk, fa);
            for (int r = 0; r < 3; ++r) {
                f[3 * a + r] = fa[r];
                for (int b = 0; b < 4; ++b)
                    for (int q = 0; q < 3; ++q) K[(3 * a + r) + 12 * (3 * b + q)] = blk[9 * b + 3 * r + q];
            }
        }
        tet_stress_out<KD>(c, p0, p1, out);
    };
    if (kind == MAT_SVK)
        run(IC<MAT_SVK>());
    else if (kind == MAT_NEOHOOKEAN)
        run(IC<MAT_NEOHOOKEAN>());
    else
        run(IC<MAT_ISOLINEAR>());
    if (!(vol > 0.0)) *A.err_flag = 1;
    for (int k = 0; k < 9; ++k) A.sig[9 * i + k] = out[k];
    const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
    for (int v = 0; v < 6; ++v) {
        A.eps[9 * i + VI[v] + 3 * VJ[v]] = out[9 + v];
        A.eps[9 * i + VJ[v] + 3 * VI[v]] = out[9 + v];
    }
}

template <int DIM>
__global__ void k_eval_trusses(EvalArgs A) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= A.count) return;
    const int64_t e = A.first + i;
    double X[2][3], U[2][3];
    for (int k = 0; k < 2; ++k) {
        const int64_t nd = A.conn[2 * e + k];
        for (int c = 0; c < DIM; ++c) {
            X[k][c] = A.X[DIM * nd + c];
            U[k][c] = A.U[DIM * nd + c];
        }
    }
    const int m = A.mat_id ? A.mat_id[e] : 0;
    const double Emod = truss_modulus(A.mat_kind[m], A.mat_params[2 * m], A.mat_params[2 * m + 1]);
    constexpr int N = 2 * DIM;
    double* K = A.K + N * N * i;
    double* f = A.f + N * i;
    double se[2];
    for (int a = 0; a < 2; ++a) {
        double blk[2][9], fa[3];
        truss_row<DIM>(A.strain_model, X, U, Emod, A.area[e], a, blk, fa, se);
        for (int r = 0; r < DIM; ++r) {
            f[DIM * a + r] = fa[r];
            for (int b = 0; b < 2; ++b)
                for (int q = 0; q < DIM; ++q) K[(DIM * a + r) + N * (DIM * b + q)] = blk[b][DIM * r + q];
        }
    }
    for (int k = 0; k < 9; ++k) {
        A.sig[9 * i + k] = 0.0;
        A.eps[9 * i + k] = 0.0;
    }
    A.sig[9 * i] = se[0];  // Trusses.jl:148-152: only [1,1] is set
    A.eps[9 * i] = se[1];
}

// ------------------------------------------------------------------------------------------------
// Jacobi-PCG
// ------------------------------------------------------------------------------------------------
// Scalars of one solve; lives in device memory, copied to pinned host memory at the end.
struct CgState {
    double rho, rho_prev, res, tol, pAp;
    double rr0, ff, uu, dd;  // ||r0||^2, ||F_ext||^2, ||U||^2 (before update), ||dU||^2
    long long it;
    int done;
    int pad;
};

enum : int { P_RR = 0, P_RZ = 1, P_FF = 2, P_UU = 3, P_PAP = 4, P_DD = 5, P_COUNT = 8 };  // 8 rows: cg_stream alternates two sets of 4

// Two-level preconditioner (precond = 2): M^-1 = D^-1 + Z E^-1 Z^T with Z = piecewise-constant translations on node
// aggregates (masked at fixed dofs) and E = Z^T K Z, inverted explicitly once per assembly (nc = BS * n_agg <= 1536).
struct CoarseArgs {
    int n_agg, nc;
    int cd;                    // coarse dofs per aggregate: BS (translations) or 6 (+ rigid-body rotations, BS = 3 only)
    int fused;                 // 1: the residual update runs in aggregate order and accumulates w = Z^T r on the way
    const double* rho;         // [n_rows][3] node position relative to its aggregate's centroid (cd = 6), else null
    const int32_t* agg;        // [n_rows] aggregate of every owned node
    const int32_t* agg_ptr;    // [n_agg+1]
    const int32_t* agg_nodes;  // [n_rows] node ids grouped by aggregate, ascending inside each
    const double* Einv;        // [nc][nc]
    double* w;                 // [nc]  Z^T r
    double* y;                 // [nc]  E^-1 w
    int n_cols;                // nodes that carry aggregate data: n_rows (rank-local level) or all local nodes (global level)
    int agg_row0;              // global id of the first aggregate of this rank (0 for the rank-local level)
    // ---- global coarse level across ranks (ONSAS_OPT_COARSE_GLOBAL, n_ranks > 1): with level-2 aggregates that are unions of
    //      this rank's aggregates (Z2 = Z T), M^-1 = D^-1 + Z [blockdiag_rank(E^-1) + T E2^-1 T^T] Z^T, E2 = Z2^T K Z2 assembled by
    //      all ranks and inverted on each.  Per application: y = E^-1 w + G w2 with G = T (E2^-1)[own rows, :] precomputed and
    //      w2 = all ranks' T^T w, all-gathered over the peer window in the LL format.
    int glob;                  // 1 = on
    int nc2, n2_own, agg2_first;   // global coarse dofs; this rank's level-2 aggregates and the global id of its first
    const double* G;               // [nc][nc2]
    const int32_t* child_ptr;      // [n2_own+1] the (consecutive) level-1 aggregates of each own level-2 aggregate
    const double* dvec;            // [n_agg][3] centroid of a level-1 aggregate minus its parent's
    unsigned long long* w2_ll;     // local LL receive buffer: [2][nc2] pairs of words
    unsigned long long* const* peer_w2;  // [n_ranks] the same buffer on every rank
    double* w2_plain;              // [nc2] the gathered vector as plain doubles (written by one CTA, read by all)
    unsigned int* w2_flag;         // epoch of the content of w2_plain
    unsigned long long* w2_epoch;  // the epoch counter, persists across launches
    int n_ranks;
};

struct CgArgs {
    int64_t n_rows;   // owned block rows
    int64_t n;        // owned dofs = n_rows * BS
    const int64_t* slice_ptr;
    const int32_t* col;
    const double* val;
    const int32_t* diag_slot;  // per row: position of the diagonal block in the row
    const uint8_t* mask;       // bit 0: free dof; bit 1: interface dof (a neighbour rank needs its value)
    double* x;                 // solution (dU), owned dofs
    double* r;
    double* p;   // search direction, n_local dofs (owned + halo)
    double* Ap;
    double* dinv;
    const double* Fext;
    const double* Fint;
    const double* rhs;  // if non-null the right-hand side is mask*rhs instead of mask*(Fext - Fint)
    double* U;          // updated in the epilogue when update_U != 0
    int update_U;
    int precond;        // 0 = none (reference default), 1 = Jacobi
    double reltol, abstol;
    long long maxiter;
    double* partials;   // [P_COUNT][part_stride]
    int part_stride;
    CgState* st;
    long long* prof;    // optional [8]: SM-clock cycles block 0 spent per phase of the persistent kernel (diagnostics)
    int* err;           // device error flag (3 = a shared-memory stage never arrived)
    double* p_pad;      // cg_stream: the vector the SpMV gathers, one 32-byte sector per node (BS = 3: stride 4), owned + halo nodes:
                        // the search direction p (classic CG) or the preconditioned residual u (single-reduction CG)
    double* s;          // single-reduction CG: s = K p (owned dofs)
    CoarseArgs co;      // precond = 2
};

// fixed-order block reduction; result valid in every thread
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) s += sh[i];
    return s;
}

// every CTA sums the per-CTA partials in the same order -> identical value everywhere
template <int NT>
__device__ __forceinline__ double sum_partials(const double* part, int nblk, double* sh) {
    double v = 0.0;
    for (int j = threadIdx.x; j < nblk; j += NT) v += __ldcg(part + j);
    return block_sum<NT>(v, sh);
}

template <int BS>
__device__ __forceinline__ double diag_entry(const CgArgs& A, int64_t i) {
    const int64_t row = i / BS;
    const int c = (int)(i % BS);
    const int64_t base = A.slice_ptr[row / C];
    const int s = A.diag_slot[row];
    return A.val[((base + s) * BS * BS + c * BS + c) * C + (row % C)];
}

// One scalar row of y = M K M p.  Work item t = (slice, block-row component r, lane): the 8 lanes of a
// (slice, r) group read one 64-byte segment of K per (block, q) -- full sectors -- and the 3 components
// of a node sit in consecutive groups, so a warp covers 32 scalar rows of at most 2 slices.
// Returns p[i] * y[i] (0 for padded rows).
template <int BS>
__device__ __forceinline__ double spmv_item(const CgArgs& A, int64_t t) {
    const int64_t sl = t / (C * BS);
    const int r = (int)((t / C) % BS);
    const int lane = (int)(t % C);
    const int64_t row = sl * C + lane;
    if (row >= A.n_rows) return 0.0;
    const int64_t base = A.slice_ptr[sl];
    const int width = (int)(A.slice_ptr[sl + 1] - base);
    const int32_t* cp = A.col + base * C + lane;
    const double* vp = A.val + (base * BS * BS + r * BS) * C + lane;
    double acc = 0.0;
#pragma unroll 4
    for (int s = 0; s < width; ++s) {
        const int64_t cnode = __ldg(cp + (int64_t)s * C);
#pragma unroll
        for (int q = 0; q < BS; ++q) acc += __ldcs(vp + ((int64_t)s * BS * BS + q) * C) * A.p[cnode * BS + q];
    }
    const int64_t i = row * BS + r;
    const double y = (A.mask[i] & 1) ? acc : 0.0;
    A.Ap[i] = y;
    return A.p[i] * y;
}

template <int BS>
__device__ __forceinline__ int64_t spmv_items(const CgArgs& A) {
    return ((A.n_rows + C - 1) / C) * (int64_t)(C * BS);
}

// ---- phase bodies shared by the persistent and the multi-launch drivers (grid-stride)
template <int BS>
__device__ __forceinline__ void cg_prologue_body(const CgArgs& A, int64_t gtid, int64_t gsz, double s[4]) {
    s[0] = s[1] = s[2] = s[3] = 0.0;
    for (int64_t i = gtid; i < A.n; i += gsz) {
        const bool m = (A.mask[i] & 1) != 0;
        double ri = 0.0;
        if (m) ri = A.rhs ? A.rhs[i] : (A.Fext[i] - A.Fint[i]);
        double di = m ? 1.0 : 0.0;
        if (m && A.precond) di = 1.0 / diag_entry<BS>(A, i);
        A.r[i] = ri;
        A.x[i] = 0.0;
        A.p[i] = 0.0;
        A.dinv[i] = di;
        const double fe = A.Fext ? A.Fext[i] : 0.0, u = A.U ? A.U[i] : 0.0;
        s[P_RR] += ri * ri;
        s[P_RZ] += ri * (ri * di);
        s[P_FF] += fe * fe;
        s[P_UU] += u * u;
    }
}

__device__ __forceinline__ void cg_update_p_body(const CgArgs& A, int64_t gtid, int64_t gsz, double beta) {
    for (int64_t i = gtid; i < A.n; i += gsz) A.p[i] = A.r[i] * A.dinv[i] + beta * A.p[i];
}

__device__ __forceinline__ void cg_update_xr_body(const CgArgs& A, int64_t gtid, int64_t gsz, double alpha, double s[2]) {
    s[0] = s[1] = 0.0;
    for (int64_t i = gtid; i < A.n; i += gsz) {
        A.x[i] += alpha * A.p[i];
        const double ri = A.r[i] - alpha * A.Ap[i];
        A.r[i] = ri;
        s[0] += ri * ri;
        s[1] += ri * (ri * A.dinv[i]);
    }
}

__device__ __forceinline__ double cg_epilogue_body(const CgArgs& A, int64_t gtid, int64_t gsz) {
    double dd = 0.0;
    for (int64_t i = gtid; i < A.n; i += gsz) {
        const double dx = A.x[i];
        dd += dx * dx;
        if (A.update_U == 1) A.U[i] += dx;  // NonLinearStaticAnalyses.jl:144 (x is zero at fixed dofs)
        else if (A.update_U == 2 && (A.mask[i] & 1)) A.U[i] = dx;  // LinearStaticAnalyses.jl:151-152: U[free] = dU
    }
    return dd;
}

constexpr int CG_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// multi-GPU persistent PCG over NVLink peer memory (one process per GPU, CUDA IPC)
// ------------------------------------------------------------------------------------------------
// Every rank runs the same persistent kernel on its own GPU.  All cross-GPU traffic is one-way P2P stores in
// the "LL" format: a double travels as two 8-byte words {32 data bits | 32-bit epoch}; an 8-byte store is atomic,
// so the receiver just polls until both words carry the expected epoch -- no fences, no separate flags, no
// acknowledgements on the critical path.
//   * scalars (p.Ap ; r.r, r.z ; the Newton norms): block 0 writes its rank's partial into a slot of every peer;
//     every CTA of every rank polls the slots and sums them in rank order -> bitwise identical everywhere;
//   * halo: the thread that updates r_i also stores z_i = r_i / d_i into the neighbours' receive buffers, i.e.
//     BEFORE the (r.r, r.z) all-reduce, so the exchange overlaps it; once beta is known each rank forms
//     p = z + beta p for its owned AND its halo nodes locally.  SpMV then reads only local memory.
// Three grid syncs per iteration, as on one GPU.  A poll that exceeds ~2 s of SM clock raises *err = 2.
constexpr int P2P_MAXR = 16;
struct P2PArgs {
    int n_ranks, rank;
    long long n_halo_dofs;
    const long long* push_ptr;               // [n_own_dofs+1] CSR: remote LL slots that want dof i
    unsigned long long* const* push_dst;     // remote addresses (16 bytes each) in the neighbours' zh buffers
    unsigned long long* zh;                  // local receive buffer: n_halo_dofs LL pairs
    unsigned long long* slots;               // local [2][P2P_MAXR][4] LL pairs
    unsigned long long* const* peer_slots;   // [n_ranks] the same array on every rank (self included)
    unsigned long long* epochs;              // local [2]: scalar epoch, halo epoch (persist across launches)
    int* err;
};

__device__ __forceinline__ void ll_store(unsigned long long* dst, double v, unsigned int flag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long f = (unsigned long long)flag << 32;
    // one 16-byte store = two atomic 8-byte words {32 data bits | 32-bit epoch}
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"((bits & 0xffffffffull) | f), "l"((bits >> 32) | f)
                 : "memory");
}

__device__ __forceinline__ double ll_load(const unsigned long long* src, unsigned int flag, int* err) {
    const volatile unsigned long long* s = src;
    unsigned long long x = s[0], y = s[1];
    if ((unsigned int)(x >> 32) != flag || (unsigned int)(y >> 32) != flag) {
        const long long t0 = clock64();
        do {
            x = s[0];
            y = s[1];
            if (clock64() - t0 > 4000000000LL) {
                *err = 2;
                break;
            }
        } while ((unsigned int)(x >> 32) != flag || (unsigned int)(y >> 32) != flag);
    }
    return __longlong_as_double((long long)((x & 0xffffffffull) | (y << 32)));
}

// all-reduce of NV <= 4 doubles across ranks; v[] holds this rank's value (identical in all its CTAs) on entry.
// NV is a compile-time constant and the values pass through shared memory, so v[] stays in registers.
template <int NV>
__device__ __forceinline__ void p2p_allreduce(const P2PArgs& P, double (&v)[NV], unsigned int& epoch, double* sh4) {
    ++epoch;
    const int par = (int)(epoch & 1u);
    const int t = threadIdx.x;
    if (blockIdx.x == 0) {
        __syncthreads();  // sh4 free for reuse
        if (t == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) sh4[k] = v[k];
        }
        __syncthreads();
        if (t < P.n_ranks * NV) {
            const int r = t / NV, k = t % NV;
            ll_store(P.peer_slots[r] + (size_t)((par * P2P_MAXR + P.rank) * 4 + k) * 2, sh4[k], epoch);
        }
    }
    __syncthreads();
    if (t < NV) {
        double s = 0.0;
        for (int r = 0; r < P.n_ranks; ++r) s += ll_load(P.slots + (size_t)((par * P2P_MAXR + r) * 4 + t) * 2, epoch, P.err);
        sh4[t] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = sh4[k];
    __syncthreads();
}

// store z_i for every peer slot that wants owned dof i
__device__ __forceinline__ void p2p_push(const P2PArgs& P, int64_t i, double z, unsigned int hepoch) {
    for (long long q = P.push_ptr[i]; q < P.push_ptr[i + 1]; ++q) ll_store(P.push_dst[q], z, hepoch);
}

// The whole linear solve of one Newton iteration in one cooperative launch, on 1 GPU or on N:
// residual r = (F_ext - F_int)[free] (StaticStates.jl:113-116), PCG exactly as IterativeSolvers'
// (P)CGIterable runs it (tolerance = max(reltol*||r0||, abstol), stop when ||r|| <= tol or
// it >= maxiter), then dU norms and U[free] += dU (NonLinearStaticAnalyses.jl:136-144).
// Multi-GPU (P.n_ranks > 0) is a RUN-TIME switch on purpose: the single- and the multi-GPU solve are the same
// machine code, so the SpMV loop's instruction schedule -- what its HBM throughput lives on -- cannot differ.
template <int BS, bool PROF, int MINB>
__global__ void __launch_bounds__(CG_THREADS, MINB) cg_persistent(CgArgs A, P2PArgs P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[CG_THREADS / 32];
    __shared__ double sh4[4];
    const int64_t gtid = blockIdx.x * (int64_t)CG_THREADS + threadIdx.x;
    const int64_t gsz = gridDim.x * (int64_t)CG_THREADS;
    const int nb = gridDim.x;
    double* part = A.partials;
    const int ps = A.part_stride;
    const bool mg = P.n_ranks > 0;
    unsigned int repoch = 0, hepoch = 0;
    if (mg) {
        repoch = (unsigned int)P.epochs[0];
        hepoch = (unsigned int)P.epochs[1];
    }

    // ---- prologue: residual, Jacobi diagonal, norms; multi-GPU: first halo push of z = r / d
    double g4[4];
    cg_prologue_body<BS>(A, gtid, gsz, g4);
    if (mg) {
        ++hepoch;
        for (int64_t i = gtid; i < A.n; i += gsz)
            if (A.mask[i] & 2) p2p_push(P, i, A.r[i] * A.dinv[i], hepoch);
        for (int64_t i = A.n + gtid; i < A.n + P.n_halo_dofs; i += gsz) {  // halo part of the search direction and of the solution
            A.p[i] = 0.0;
            A.x[i] = 0.0;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double b = block_sum<CG_THREADS>(g4[k], sh);
        if (threadIdx.x == 0) part[k * ps + blockIdx.x] = b;
    }
    grid.sync();
#pragma unroll
    for (int k = 0; k < 4; ++k) g4[k] = sum_partials<CG_THREADS>(part + k * ps, nb, sh);
    if (mg) p2p_allreduce<4>(P, g4, repoch, sh4);
    const double rr0 = g4[P_RR], ff = g4[P_FF], uu = g4[P_UU];
    double rho = g4[P_RZ];
    double res = sqrt(rr0);
    const double tol = fmax(A.reltol * res, A.abstol);
    double rho_prev = 1.0;
    long long it = 0;

    const bool profiling = PROF && A.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long tprev = profiling ? clock64() : 0;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define CG_PROF(k)                          \
    if constexpr (PROF) {                   \
        if (profiling) {                    \
            const long long tn = clock64(); \
            tacc[k] += tn - tprev;          \
            tprev = tn;                     \
        }                                   \
    }
    while (!(it >= A.maxiter || res <= tol || res != res)) {  // a non-finite residual (breakdown) ends the solve: err = 4
        const double beta = rho / rho_prev;
        // ---- p = z + beta p on owned dofs; on halo dofs z comes from the neighbours' pushes of epoch hepoch
        cg_update_p_body(A, gtid, gsz, beta);
        if (mg) {
            for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz)
                A.p[A.n + h] = ll_load(P.zh + 2 * h, hepoch, P.err) + beta * A.p[A.n + h];
        }
        CG_PROF(0)
        grid.sync();
        CG_PROF(1)
        long long tc0 = 0;
        if constexpr (PROF) tc0 = clock64();
        double d = 0.0;
        for (int64_t t = gtid, nt = spmv_items<BS>(A); t < nt; t += gsz) d += spmv_item<BS>(A, t);
        d = block_sum<CG_THREADS>(d, sh);
        if (threadIdx.x == 0) part[P_PAP * ps + blockIdx.x] = d;
        if constexpr (PROF) {
            if (A.prof != nullptr && threadIdx.x == 0) {  // per-CTA SpMV cycles and the SM it runs on (diagnostics)
                A.prof[16 + blockIdx.x] += clock64() - tc0;
                unsigned int smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                A.prof[16 + 2048 + blockIdx.x] = smid;
            }
        }
        CG_PROF(2)
        grid.sync();
        CG_PROF(3)
        double pAp[1] = {sum_partials<CG_THREADS>(part + P_PAP * ps, nb, sh)};
        if (mg) p2p_allreduce<1>(P, pAp, repoch, sh4);
        const double alpha = rho / pAp[0];
        // ---- x += alpha p ; r -= alpha Ap ; multi-GPU: push z of the interface dofs right away, so the exchange
        //      overlaps the (r.r, r.z) all-reduce
        ++hepoch;
        double s2[2] = {0.0, 0.0};
        for (int64_t i = gtid; i < A.n; i += gsz) {
            A.x[i] += alpha * A.p[i];
            const double ri = A.r[i] - alpha * A.Ap[i];
            A.r[i] = ri;
            const double zi = ri * A.dinv[i];
            s2[0] += ri * ri;
            s2[1] += ri * zi;
            if (mg && (A.mask[i] & 2)) p2p_push(P, i, zi, hepoch);
        }
        if (mg)  // the solution on the halo dofs follows from the halo search directions: no exchange of U after the solve
            for (int64_t i = A.n + gtid; i < A.n + P.n_halo_dofs; i += gsz) A.x[i] += alpha * A.p[i];
        const double b0 = block_sum<CG_THREADS>(s2[0], sh);
        const double b1 = block_sum<CG_THREADS>(s2[1], sh);
        if (threadIdx.x == 0) {
            part[P_RR * ps + blockIdx.x] = b0;
            part[P_RZ * ps + blockIdx.x] = b1;
        }
        CG_PROF(4)
        grid.sync();
        CG_PROF(5)
        s2[0] = sum_partials<CG_THREADS>(part + P_RR * ps, nb, sh);
        s2[1] = sum_partials<CG_THREADS>(part + P_RZ * ps, nb, sh);
        if (mg) p2p_allreduce<2>(P, s2, repoch, sh4);
        rho_prev = rho;
        rho = s2[1];
        res = sqrt(s2[0]);
        ++it;
        CG_PROF(6)
    }
#undef CG_PROF
    if constexpr (PROF) {
        if (profiling)
            for (int k = 0; k < 8; ++k) A.prof[k] = tacc[k];
    }

    double dd = cg_epilogue_body(A, gtid, gsz);
    if (mg && A.update_U)  // halo dofs: the same update as on their owner (the mask carries the free bit of halo dofs too)
        for (int64_t i = A.n + gtid; i < A.n + P.n_halo_dofs; i += gsz) {
            if (A.update_U == 1) A.U[i] += A.x[i];
            else if (A.mask[i] & 1) A.U[i] = A.x[i];
        }
    dd = block_sum<CG_THREADS>(dd, sh);
    if (threadIdx.x == 0) part[P_DD * ps + blockIdx.x] = dd;
    grid.sync();
    double dd1[1] = {sum_partials<CG_THREADS>(part + P_DD * ps, nb, sh)};
    if (mg) p2p_allreduce<1>(P, dd1, repoch, sh4);
    dd = dd1[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        CgState* st = A.st;
        st->rho = rho;
        st->rho_prev = rho_prev;
        st->res = res;
        st->tol = tol;
        st->pAp = 0.0;
        st->rr0 = rr0;
        st->ff = ff;
        st->uu = uu;
        st->dd = dd;
        st->it = it;
        st->done = 1;
        if (res != res) *A.err = 4;
        if (mg) {
            P.epochs[0] = repoch;
            P.epochs[1] = hepoch;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// two-level preconditioner: coarse operator E = Z^T (M K M) Z and its explicit inverse
// ------------------------------------------------------------------------------------------------
// One CTA per aggregate a builds the BS rows 3a..3a+2 of E in shared memory.  Its nodes are visited in ascending order
// and a node's blocks in storage order by the same BS*BS threads, so every entry of E is summed in a fixed order
// (deterministic); the loads of a row (column ids -> aggregate ids, values) are issued by the whole CTA in parallel.
constexpr int CO_THREADS = 256;
// (the warps-per-aggregate split of the w = Z^T r reduction is chosen at run time inside cg_stream)
constexpr int CO_NW = CO_THREADS / 32;  // nodes of the aggregate in flight (one per warp)
constexpr int CO_WB = 32;               // blocks of a node staged per round
// entry (p, k) of R = -skew(rho): the displacement of a node at rho under a unit rotation about axis k (u = omega x rho)
__device__ __forceinline__ double rot_entry(const double* rho, int p, int k) {
    if (p == k) return 0.0;
    const double v = rho[3 - p - k];
    return (k - p + 3) % 3 == 1 ? v : -v;
}
// column `cdof` of Z_i = [I | R_i] (or [I] without rotations), rows of fixed dofs zeroed
template <int BS>
__device__ __forceinline__ void z_column(int cdof, const double* rho, unsigned mask, double (&z)[BS]) {
#pragma unroll
    for (int p = 0; p < BS; ++p) {
        double v = cdof < BS ? (p == cdof ? 1.0 : 0.0) : rot_entry(rho, p, cdof - BS);
        z[p] = ((mask >> p) & 1u) ? v : 0.0;
    }
}
// The first form of the coarse assembly (every contribution applied by the same CD*CD threads, one after the other): kept as the
// cross-check of k_coarse_assemble (ONSAS_COARSE_CHECK) and as ONSAS_COARSE_SERIAL = 1.
template <int BS>
__global__ void __launch_bounds__(CO_THREADS) k_coarse_assemble_serial(CgArgs A, double* E) {
    extern __shared__ double erow[];  // [CD][nc]
    constexpr int BB = BS * BS;
    const CoarseArgs& G = A.co;
    const int a = blockIdx.x, tid = threadIdx.x, nc = G.nc, CD = G.cd, warp = tid >> 5, lane = tid & 31;
    const bool rbm = CD > BS;
    // the staged block values are a separate (static) array: the apply loop below reads them while it read-modify-writes
    // erow, and only distinct objects let the compiler hoist those reads over the stores
    __shared__ double sval[CO_NW * CO_WB * BB];
    __shared__ int s_b[CO_NW][CO_WB];            // aggregate of the block's column (-1: halo column)
    __shared__ unsigned char s_m[CO_NW][CO_WB];  // mask bits of the column node's dofs
    __shared__ double s_rj[CO_NW][CO_WB][3];     // rho of the column node
    __shared__ double s_ri[CO_NW][3];            // rho of the row node
    __shared__ int s_nb[CO_NW];                  // blocks the warp staged this round
    __shared__ unsigned s_mi[CO_NW];             // mask bits of the row node's dofs
    for (int k = tid; k < CD * nc; k += CO_THREADS) erow[k] = 0.0;
    __syncthreads();
    // Eight nodes are fetched at a time, one per warp (the four dependent load levels node id -> slice -> column ids
    // -> aggregate ids overlap across the warps); CD*CD threads then apply the staged blocks in (node, block) order:
    // E[a-rows, b-cols] += Z_i^T (M K_ij M) Z_j.
    const int q0 = G.agg_ptr[a], q1 = G.agg_ptr[a + 1];
    for (int qb = q0; qb < q1; qb += CO_NW) {
        const int q = qb + warp;
        int64_t base = 0, inode = 0;
        int width = 0, lrow = 0;
        unsigned mi = 0;
        if (q < q1) {
            inode = G.agg_nodes[q];
            base = A.slice_ptr[inode / C];
            width = (int)(A.slice_ptr[inode / C + 1] - base);
            lrow = (int)(inode % C);
            for (int r = 0; r < BS; ++r) mi |= (unsigned)(A.mask[inode * BS + r] & 1) << r;
        }
        for (int s0 = 0;; s0 += CO_WB) {
            const int left = width - s0;
            const int nb = left <= 0 ? 0 : (left < CO_WB ? left : CO_WB);
            if (lane < nb) {
                const int64_t j = A.col[(base + s0 + lane) * C + lrow];
                int b = -1;
                unsigned mj = 0;
                if (j < G.n_cols) {
                    b = G.agg[j];
                    for (int c = 0; c < BS; ++c) mj |= (unsigned)(A.mask[j * BS + c] & 1) << c;
                    if (rbm)
                        for (int c = 0; c < 3; ++c) s_rj[warp][lane][c] = G.rho[j * 3 + c];
                }
                s_b[warp][lane] = b;
                s_m[warp][lane] = (unsigned char)mj;
            }
            for (int t = lane; t < nb * BB; t += 32) sval[(size_t)warp * CO_WB * BB + t] = A.val[((base + s0 + t / BB) * BB + t % BB) * C + lrow];
            if (lane == 0) {
                s_nb[warp] = nb;
                s_mi[warp] = mi;
            }
            if (rbm && lane < 3 && nb > 0) s_ri[warp][lane] = G.rho[inode * 3 + lane];
            __syncthreads();
            if (tid < CD * CD) {  // thread (r, c): nodes in ascending order, blocks in storage order -> a fixed summation order
                const int r = tid / CD, c = tid % CD;
                // consecutive blocks mostly belong to the same target aggregate (a node's neighbours are its own aggregate's
                // nodes, except at the aggregate's surface): their contributions are summed in a register and meet the row in
                // shared memory once per run -- no read-modify-write chain through shared memory per block
                int run_b = -1;
                double run = 0.0;
                for (int w = 0; w < CO_NW; ++w) {
                    const int nbw = s_nb[w];
                    if (nbw == 0) continue;
                    double zi[BS];
                    z_column<BS>(r, s_ri[w], s_mi[w], zi);
                    for (int s = 0; s < nbw; ++s) {
                        const int b = s_b[w][s];
                        if (b < 0) continue;
                        double zj[BS];
                        z_column<BS>(c, s_rj[w][s], s_m[w][s], zj);
                        const double* Kb = sval + ((size_t)w * CO_WB + s) * BB;
                        double acc = 0.0;
#pragma unroll
                        for (int p = 0; p < BS; ++p) {
                            double t = 0.0;
#pragma unroll
                            for (int qq = 0; qq < BS; ++qq) t += Kb[p * BS + qq] * zj[qq];
                            acc += zi[p] * t;
                        }
                        if (b != run_b) {
                            if (run_b >= 0) erow[(size_t)r * nc + run_b * CD + c] += run;
                            run_b = b;
                            run = 0.0;
                        }
                        run += acc;
                    }
                }
                if (run_b >= 0) erow[(size_t)r * nc + run_b * CD + c] += run;
            }
            if (!__syncthreads_or(left > CO_WB)) break;  // barrier (staging buffers are free again) + "another round?"
        }
    }
    // coarse dofs without any free fine dof: unit diagonal keeps E invertible (their w is always 0).  With rotations a
    // degenerate aggregate (collinear nodes) has a rotation that moves nothing: E + 1e-9 diag(E) stays positive definite.
    if (tid < CD) {
        double& d = erow[(size_t)tid * nc + (a + G.agg_row0) * CD + tid];
        if (d == 0.0) d = 1.0;
        else if (rbm) d *= 1.0 + 1e-9;
    }
    __syncthreads();
    for (int k = tid; k < CD * nc; k += CO_THREADS) E[(size_t)(a * CD) * nc + k] = erow[k];
}

// E[a-rows, :] = sum over the nodes i of aggregate a and the blocks (i, j) of their rows of Z_i^T (M K_ij M) Z_j.
// One CTA per aggregate; eight of its nodes are fetched at a time, one per warp (the four dependent load levels node id ->
// slice -> column ids -> aggregate ids overlap across the warps).  Most blocks of an aggregate's rows have their column in the
// SAME aggregate (all but the surface nodes' outward neighbours), so the serial form above spent its time in one dependent
// chain of ~10^4 contributions to E[a, a] by 36 threads, the other 220 waiting at the barrier (ncu: 15.8 barrier stalls per
// issued instruction, 3.06 ms for 256 aggregates).  Here
//  * E[a, a] is accumulated by ALL warps: each warp sums the own-aggregate blocks of the nodes it staged in registers (one
//    lane per entry of the upper triangle -- the sum over all (i, j) pairs inside an aggregate is symmetric because K is),
//    and the eight partial sums meet once, in warp order, at the end;
//  * the blocks that target OTHER aggregates (compacted into a list at staging time) are dealt by target over
//    CO_THREADS / (CD*CD) thread groups (target % groups), each applying its targets in (node, block) order as before.
// Every entry of E still has one fixed summation order.
template <int BS>
__global__ void __launch_bounds__(CO_THREADS) k_coarse_assemble(CgArgs A, double* E) {
    extern __shared__ double erow[];  // [CD][nc]
    constexpr int BB = BS * BS;
    constexpr int CDX = BS == 3 ? 6 : BS;  // largest CD
    const CoarseArgs& G = A.co;
    const int a = blockIdx.x, tid = threadIdx.x, nc = G.nc, CD = G.cd, warp = tid >> 5, lane = tid & 31;
    const bool rbm = CD > BS;
    const int b_own = a + G.agg_row0;
    __shared__ double sval[CO_NW * CO_WB * BB];
    __shared__ int s_b[CO_NW][CO_WB];            // aggregate of the block's column (-1: halo column)
    __shared__ unsigned char s_m[CO_NW][CO_WB];  // mask bits of the column node's dofs
    __shared__ double s_rj[CO_NW][CO_WB][3];     // rho of the column node
    __shared__ double s_ri[CO_NW][3];            // rho of the row node
    __shared__ int s_nb[CO_NW];                  // blocks the warp staged this round
    __shared__ unsigned s_mi[CO_NW];             // mask bits of the row node's dofs
    __shared__ unsigned char s_list[CO_NW][CO_WB];  // staged blocks of the warp whose column lies in ANOTHER aggregate
    __shared__ unsigned char s_grp[CO_NW][CO_WB];   // ... and the thread group that owns that aggregate's columns of E
    __shared__ int s_cnt[CO_NW];
    __shared__ double s_own[CO_NW][CDX * (CDX + 1) / 2];
    for (int k = tid; k < CD * nc; k += CO_THREADS) erow[k] = 0.0;
    // lane -> entry (pr, pc), pr <= pc, of the upper triangle of E[a, a]
    const int NU = CD * (CD + 1) / 2;
    int pr = 0, pc = lane;
    while (pr < CD && pc >= CD - pr) {
        pc -= CD - pr;
        ++pr;
    }
    pc += pr;
    const bool p_act = lane < NU;
    double own = 0.0;
    // thread -> (group, entry) for the other aggregates' blocks
    const int NG = CO_THREADS / (CD * CD);
    const int grp = tid / (CD * CD), er = (tid % (CD * CD)) / CD, ec = tid % CD;
    __syncthreads();
    const int q0 = G.agg_ptr[a], q1 = G.agg_ptr[a + 1];
    for (int qb = q0; qb < q1; qb += CO_NW) {
        const int q = qb + warp;
        int64_t base = 0, inode = 0;
        int width = 0, lrow = 0;
        unsigned mi = 0;
        if (q < q1) {
            inode = G.agg_nodes[q];
            base = A.slice_ptr[inode / C];
            width = (int)(A.slice_ptr[inode / C + 1] - base);
            lrow = (int)(inode % C);
            for (int r = 0; r < BS; ++r) mi |= (unsigned)(A.mask[inode * BS + r] & 1) << r;
        }
        for (int s0 = 0;; s0 += CO_WB) {
            const int left = width - s0;
            const int nb = left <= 0 ? 0 : (left < CO_WB ? left : CO_WB);
            int b = -1;
            if (lane < nb) {
                const int64_t j = A.col[(base + s0 + lane) * C + lrow];
                unsigned mj = 0;
                if (j < G.n_cols) {
                    b = G.agg[j];
                    for (int c = 0; c < BS; ++c) mj |= (unsigned)(A.mask[j * BS + c] & 1) << c;
                    if (rbm)
                        for (int c = 0; c < 3; ++c) s_rj[warp][lane][c] = G.rho[j * 3 + c];
                }
                s_b[warp][lane] = b;
                s_m[warp][lane] = (unsigned char)mj;
            }
            const bool other = b >= 0 && b != b_own;
            const unsigned om = __ballot_sync(0xffffffffu, other);
            if (other) {
                const int pos = __popc(om & ((1u << lane) - 1u));
                s_list[warp][pos] = (unsigned char)lane;
                s_grp[warp][pos] = (unsigned char)(b % NG);
            }
            for (int t = lane; t < nb * BB; t += 32) sval[(size_t)warp * CO_WB * BB + t] = A.val[((base + s0 + t / BB) * BB + t % BB) * C + lrow];
            if (lane == 0) {
                s_nb[warp] = nb;
                s_mi[warp] = mi;
                s_cnt[warp] = __popc(om);
            }
            if (rbm && lane < 3 && nb > 0) s_ri[warp][lane] = G.rho[inode * 3 + lane];
            __syncthreads();
            // ---- E[a, a]: this warp's node, blocks in storage order
            if (p_act && nb > 0) {
                double zi[BS];
                z_column<BS>(pr, s_ri[warp], mi, zi);
                // (branch-free and unrolled: the blocks' loads and products overlap, only the additions into `own` are a chain;
                // a block of another aggregate contributes an exact zero through its cleared mask)
#pragma unroll 4
                for (int s = 0; s < nb; ++s) {
                    const unsigned mj = s_b[warp][s] == b_own ? (unsigned)s_m[warp][s] : 0u;
                    double zj[BS];
                    z_column<BS>(pc, s_rj[warp][s], mj, zj);
                    const double* Kb = sval + ((size_t)warp * CO_WB + s) * BB;
                    double acc = 0.0;
#pragma unroll
                    for (int p = 0; p < BS; ++p) {
                        double t = 0.0;
#pragma unroll
                        for (int qq = 0; qq < BS; ++qq) t += Kb[p * BS + qq] * zj[qq];
                        acc += zi[p] * t;
                    }
                    own += acc;
                }
            }
            // ---- the other aggregates: group grp owns the targets b with b % NG == grp; nodes in ascending order, blocks in
            //      storage order; consecutive contributions to the same target are summed in a register first
            if (grp < NG) {
                int run_b = -1;
                double run = 0.0;
                for (int w = 0; w < CO_NW; ++w) {
                    const int cnt = s_cnt[w];
                    bool have_zi = false;
                    double zi[BS];
                    for (int k = 0; k < cnt; ++k) {
                        if (s_grp[w][k] != grp) continue;
                        const int s = s_list[w][k];
                        const int bt = s_b[w][s];
                        if (!have_zi) {
                            z_column<BS>(er, s_ri[w], s_mi[w], zi);
                            have_zi = true;
                        }
                        double zj[BS];
                        z_column<BS>(ec, s_rj[w][s], s_m[w][s], zj);
                        const double* Kb = sval + ((size_t)w * CO_WB + s) * BB;
                        double acc = 0.0;
#pragma unroll
                        for (int p = 0; p < BS; ++p) {
                            double t = 0.0;
#pragma unroll
                            for (int qq = 0; qq < BS; ++qq) t += Kb[p * BS + qq] * zj[qq];
                            acc += zi[p] * t;
                        }
                        if (bt != run_b) {
                            if (run_b >= 0) erow[(size_t)er * nc + run_b * CD + ec] += run;
                            run_b = bt;
                            run = 0.0;
                        }
                        run += acc;
                    }
                }
                if (run_b >= 0) erow[(size_t)er * nc + run_b * CD + ec] += run;
            }
            if (!__syncthreads_or(left > CO_WB)) break;  // barrier (staging buffers are free again) + "another round?"
        }
    }
    // the warps' partial sums of E[a, a], in warp order; both triangles get the same value
    if (p_act) s_own[warp][lane] = own;
    __syncthreads();
    if (tid < NU) {
        double v = 0.0;
        for (int w = 0; w < CO_NW; ++w) v += s_own[w][tid];
        erow[(size_t)pr * nc + b_own * CD + pc] = v;
        erow[(size_t)pc * nc + b_own * CD + pr] = v;
    }
    __syncthreads();
    // coarse dofs without any free fine dof: unit diagonal keeps E invertible (their w is always 0).  With rotations a
    // degenerate aggregate (collinear nodes) has a rotation that moves nothing: E + 1e-9 diag(E) stays positive definite.
    if (tid < CD) {
        double& d = erow[(size_t)tid * nc + b_own * CD + tid];
        if (d == 0.0) d = 1.0;
        else if (rbm) d *= 1.0 + 1e-9;
    }
    __syncthreads();
    for (int k = tid; k < CD * nc; k += CO_THREADS) E[(size_t)(a * CD) * nc + k] = erow[k];
}

// ---- global coarse level: exchange of the rows of E2 over the peer window, and the folded operator G
// copies this rank's rows of E2 into every peer's E2 (plain peer stores; the flags follow in a later kernel of the same stream)
__global__ void k_push_rows(const double* __restrict__ src, size_t n, double* const* peers, int n_ranks, int rank, size_t dst_off) {
    for (int r = 0; r < n_ranks; ++r) {
        if (r == rank) continue;
        double* dst = peers[r] + dst_off;
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
}
// "my rows of refresh `epoch` have landed": one flag per source rank in every peer's window
__global__ void k_set_flags(unsigned long long* const* peer_flags, int n_ranks, int rank, unsigned long long epoch) {
    const int r = threadIdx.x;
    if (r < n_ranks && r != rank) {
        __threadfence_system();
        *(volatile unsigned long long*)(peer_flags[r] + rank) = epoch;
    }
}
// waits until every other rank's rows of refresh `epoch` have landed here (watchdog: err = 2)
__global__ void k_wait_flags(const unsigned long long* flags, int n_ranks, int rank, unsigned long long epoch, int* err) {
    const int r = threadIdx.x;
    if (r < n_ranks && r != rank) {
        const long long t0 = clock64();
        while (*(volatile const unsigned long long*)(flags + r) < epoch) {
            if (clock64() - t0 > 8000000000LL) {
                *err = 2;
                break;
            }
        }
    }
    __threadfence_system();
}
// G = T (E2^-1)[own rows, :]: row (a, i) of G is the row of E2^-1 of a's parent for the same coarse dof, plus, for the
// translations, the parent's rotation rows times the offset d = c(a) - c(parent)  (t_child = t_parent + omega_parent x d)
__global__ void k_build_G(const double* __restrict__ E2inv, int nc2, int nc, const int32_t* __restrict__ parent_of, int agg2_first,
                          const double* __restrict__ dvec, double* __restrict__ G) {
    const int k = blockIdx.x;
    if (k >= nc) return;
    const int a = k / 6, i = k % 6;
    const size_t gp = (size_t)(agg2_first + parent_of[a]) * 6;
    const double d0 = dvec[a * 3], d1 = dvec[a * 3 + 1], d2 = dvec[a * 3 + 2];
    for (int j = threadIdx.x; j < nc2; j += blockDim.x) {
        double v = E2inv[(gp + i) * nc2 + j];
        if (i == 0) v += d2 * E2inv[(gp + 4) * nc2 + j] - d1 * E2inv[(gp + 5) * nc2 + j];
        else if (i == 1) v += d0 * E2inv[(gp + 5) * nc2 + j] - d2 * E2inv[(gp + 3) * nc2 + j];
        else if (i == 2) v += d1 * E2inv[(gp + 3) * nc2 + j] - d0 * E2inv[(gp + 4) * nc2 + j];
        G[(size_t)k * nc2 + j] = v;
    }
}

// In-place Gauss-Jordan inversion of the SPD coarse matrix (no pivoting) in ONE cooperative launch.  The matrix is
// distributed by rows over the CTAs' shared memory (<= 11 rows of 1536 doubles = 135 KB per CTA); per elimination
// step only the pivot row travels (its owner publishes the already updated row k+1 before the step's single grid
// barrier: look-ahead), so a step costs one barrier + 12 KB of L2 reads per CTA instead of a pass over the matrix.
constexpr int GJ_THREADS = 512;
constexpr int GJ_MAX_ROWS = 12;  // rows of the matrix per CTA (shared memory: 12 x 1536 doubles = 144 KB)
__global__ void __launch_bounds__(GJ_THREADS, 1) k_gj_invert(double* M, int nc, int rows_per_cta, double* rowbuf /*[2][nc]*/) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sm[];  // [rows_per_cta][nc]
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * rows_per_cta;
    const int nr = r0 >= nc ? 0 : (nc - r0 < rows_per_cta ? nc - r0 : rows_per_cta);
    for (int k = tid; k < nr * nc; k += GJ_THREADS) sm[k] = M[(size_t)r0 * nc + k];
    __syncthreads();
    if (r0 == 0)
        for (int j = tid; j < nc; j += GJ_THREADS) rowbuf[j] = sm[j];
    grid.sync();
    for (int k = 0; k < nc; ++k) {
        const double* prow = rowbuf + (size_t)(k & 1) * nc;   // row k as it was before this step
        double* nrow = rowbuf + (size_t)((k + 1) & 1) * nc;   // row k+1 after this step
        // column k of the CTA's rows before anybody overwrites it (a thread owns whole columns below)
        double mik[GJ_MAX_ROWS];
#pragma unroll
        for (int li = 0; li < GJ_MAX_ROWS; ++li) mik[li] = li < nr ? sm[li * nc + k] : 0.0;
        const double ip = 1.0 / __ldcg(prow + k);
        __syncthreads();
        for (int j = tid; j < nc; j += GJ_THREADS) {
            const double pj = __ldcg(prow + j);
#pragma unroll
            for (int li = 0; li < GJ_MAX_ROWS; ++li) {
                if (li < nr) {
                    const int i = r0 + li;
                    double v;
                    if (i == k) v = j == k ? ip : pj * ip;
                    else v = j == k ? -mik[li] * ip : sm[li * nc + j] - mik[li] * pj * ip;
                    sm[li * nc + j] = v;
                    if (i == k + 1) nrow[j] = v;
                }
            }
        }
        grid.sync();
    }
    for (int k = tid; k < nr * nc; k += GJ_THREADS) M[(size_t)r0 * nc + k] = sm[k];
}

// Blocked form of the same inversion (default): the rows of one CTA (GJ_MAX_ROWS = 12) are one pivot PANEL, so the
// matrix is eliminated in nc / 12 steps instead of nc -- 128 grid barriers instead of 1536 for the largest coarse space.
// Step k, with P the panel's rows / columns, D = A[P,P] and R the panel's rows after its own factorisation
// (R = D^-1 A[P,:], R[:,P] = D^-1; published to `rowbuf` by its owner):
//     every other CTA:  L = A[rows,P];  A[rows,j] = (j in P ? 0 : A[rows,j]) - L R[:,j]
// and the owner of panel k+1, as soon as its rows are updated, factorises and publishes them before the step's
// single grid barrier (look-ahead), so the critical path of a step is update + 12x12 inverse + one barrier.
// Same fixed operation order on every run: the explicit inverse stays bitwise reproducible.
constexpr int GJ_B = GJ_MAX_ROWS;
__global__ void __launch_bounds__(GJ_THREADS, 1) k_gj_invert_blocked(double* M, int nc, double* rowbuf /*[2][GJ_B][nc]*/) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sm[];      // [GJ_B][nc] rows of this CTA, then L [GJ_B][GJ_B], D [GJ_B][GJ_B]
    double* Lm = sm + (size_t)GJ_B * nc;
    double* Dm = Lm + GJ_B * GJ_B;
    const int tid = threadIdx.x, c = blockIdx.x;
    const int r0 = c * GJ_B;
    const int nr = nc - r0 < GJ_B ? nc - r0 : GJ_B;  // >= 1: the grid is ceil(nc / GJ_B)
    const int np = (nc + GJ_B - 1) / GJ_B;
    for (int k = tid; k < nr * nc; k += GJ_THREADS) sm[k] = M[(size_t)r0 * nc + k];
    __syncthreads();

    // this CTA's rows are the pivot panel: D^-1 by unblocked Gauss-Jordan in shared memory (no pivoting: SPD), then
    // rows <- D^-1 rows with D^-1 itself in the panel's columns; the result is stored locally and published
    auto factor_and_publish = [&](double* out) {
        const int i = tid / GJ_B, j = tid % GJ_B;
        const bool mine = tid < GJ_B * GJ_B && i < nr && j < nr;
        if (mine) Dm[i * GJ_B + j] = sm[i * nc + r0 + j];
        __syncthreads();
        for (int s = 0; s < nr; ++s) {
            double v = 0.0;
            if (mine) {
                const double ip = 1.0 / Dm[s * GJ_B + s];
                const double dis = Dm[i * GJ_B + s], dsj = Dm[s * GJ_B + j];
                if (i == s) v = j == s ? ip : dsj * ip;
                else v = j == s ? -dis * ip : Dm[i * GJ_B + j] - dis * dsj * ip;
            }
            __syncthreads();
            if (mine) Dm[i * GJ_B + j] = v;
            __syncthreads();
        }
        for (int col = tid; col < nc; col += GJ_THREADS) {
            double a[GJ_B];
#pragma unroll
            for (int q = 0; q < GJ_B; ++q) a[q] = q < nr ? sm[q * nc + col] : 0.0;
            const bool inP = col >= r0 && col < r0 + nr;
#pragma unroll
            for (int li = 0; li < GJ_B; ++li) {
                if (li < nr) {
                    double v;
                    if (inP) {
                        v = Dm[li * GJ_B + (col - r0)];
                    } else {
                        v = 0.0;
#pragma unroll
                        for (int q = 0; q < GJ_B; ++q)
                            if (q < nr) v += Dm[li * GJ_B + q] * a[q];
                    }
                    sm[li * nc + col] = v;
                    out[(size_t)li * nc + col] = v;
                }
            }
        }
    };

    if (c == 0) factor_and_publish(rowbuf);
    grid.sync();
    for (int k = 0; k < np; ++k) {
        const int pk0 = k * GJ_B;
        const int pn = nc - pk0 < GJ_B ? nc - pk0 : GJ_B;
        const double* R = rowbuf + (size_t)(k & 1) * GJ_B * nc;
        if (c != k) {
            for (int t = tid; t < GJ_B * GJ_B; t += GJ_THREADS) {
                const int li = t / GJ_B, q = t % GJ_B;
                Lm[t] = (li < nr && q < pn) ? sm[li * nc + pk0 + q] : 0.0;  // zero padding: the loops below run to GJ_B
            }
            __syncthreads();
            for (int col = tid; col < nc; col += GJ_THREADS) {
                double rj[GJ_B];
#pragma unroll
                for (int q = 0; q < GJ_B; ++q) rj[q] = q < pn ? __ldcg(R + (size_t)q * nc + col) : 0.0;
                const bool inP = col >= pk0 && col < pk0 + pn;
#pragma unroll
                for (int li = 0; li < GJ_B; ++li) {
                    if (li < nr) {
                        double acc = inP ? 0.0 : sm[li * nc + col];
#pragma unroll
                        for (int q = 0; q < GJ_B; ++q) acc -= Lm[li * GJ_B + q] * rj[q];
                        sm[li * nc + col] = acc;
                    }
                }
            }
            __syncthreads();
            if (c == k + 1) factor_and_publish(rowbuf + (size_t)((k + 1) & 1) * GJ_B * nc);
        }
        grid.sync();
    }
    for (int k = tid; k < nr * nc; k += GJ_THREADS) M[(size_t)r0 * nc + k] = sm[k];
}

// ------------------------------------------------------------------------------------------------
// cg_stream: the persistent PCG with K streamed through shared memory by the TMA engine (cp.async.bulk)
// ------------------------------------------------------------------------------------------------
// Why: the register-fed SpMV above keeps only a handful of 8-byte loads in flight per thread, so its HBM
// throughput hangs on occupancy and on how ptxas happens to schedule the loop (measured: 48..115 us per CG
// iteration for functionally identical variants).  Here ONE CTA per SM runs N_CW consumer warps and one producer
// warp.  A consumer warp owns a private ring of DEPTH shared-memory slots; a slot holds the contiguous BSELL image
// of one slice (values + column ids, 9.1 KB on the structured tet mesh).  Lane w of the producer warp feeds
// consumer warp w: it waits on the slot's "empty" mbarrier, arms the "full" mbarrier with the byte count and
// issues two bulk copies (UBLKCP) that complete on it.  ~110-150 KB per SM are in flight whatever the compute
// warps do, and no register is spent on it.  Consumers never meet at a block barrier inside the SpMV: each waits
// for its own slice, reads K from shared memory (conflict-free: lane = (row of the slice, quarter of its blocks)),
// gathers p with one 32-byte load per block column from a sector-padded copy of p, reduces the four partial row
// sums with two xor-shuffles and releases the slot.
//   * slices are dealt round-robin (slice = (cta + k*grid) * N_CW + warp), so every CTA sweeps all of K's range;
//   * the rings keep running ACROSS CG iterations: the producer is always DEPTH fills ahead, so the first slices
//     of the next SpMV land while the vector updates and reductions of this iteration execute;
//   * K's lines carry an L2 evict_first hint: the CG vectors, which are re-read within an iteration, stay in L2;
//   * a row's blocks are summed in 4 interleaved partial sums -- a fixed order, so the solve stays bitwise
//     reproducible (it differs in rounding from the 1-thread-per-row order of the other drivers);
//   * multi-GPU (P.n_ranks > 0) is a run-time switch: same machine code on 1 and on N GPUs.
struct StreamArgs {
    int n_cw;         // consumer warps (<= ST_MAX_CW)
    int depth;        // ring slots per consumer warp (2..4)
    int slot_blocks;  // capacity of a slot in blocks (widest slice)
    long long n_slices;
    int l2_prefetch;  // slices per consumer warp the producer pulls into L2 behind the ring while HBM idles (vector phases, barriers)
};

constexpr int ST_MAX_CW = 12;
constexpr int ST_MAX_DEPTH = 4;


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait with a watchdog: a lost copy / a stuck consumer raises *err = 3 instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
    if (mbar_try_wait(bar, parity)) return;
    if (*(volatile int*)err == 3) return;  // already failing: do not wait again
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000LL) {
            *(volatile int*)err = 3;
            return;
        }
    }
}
// K is read once per CG iteration and is larger than L2: its lines are marked evict_first so that the CG vectors
// (which ARE re-read within an iteration) keep their place in L2
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
// HBM idles while the vector phases and the grid barriers of an iteration run out of L2: the producer uses that window to
// pull the slices that FOLLOW the ring's content into L2, so the first part of the next SpMV phase streams at L2 speed
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// p of one node from the padded search direction: a single sector, a single load instruction
template <int BS>
__device__ __forceinline__ void ld_node(const double* pn, double (&v)[BS]) {
    if constexpr (BS == 3) {
        double pad;
        asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(pad) : "l"(pn));
    } else if constexpr (BS == 2) {
        asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(pn));
    } else {
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(v[0]) : "l"(pn));
    }
}

// CW = consumer warps the instantiation is compiled for (8: 288 threads, 224 registers; 12: 416 threads, 152 registers)
// SR = single-reduction CG (Chronopoulos-Gear), for precond 0 / 1: with u = M^-1 r, w = K u, gamma = r.u, delta = u.w
//     beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old),
//     p = u + beta p,  s = w + beta s (= K p),  x += alpha p,  r -= alpha s,  u = M^-1 r
// All vector updates are ONE local phase, and r.r, r.u are known at its end, so an iteration is
// {vector phase | grid barrier | SpMV w = K u with u.w | one reduction of (r.r, r.u, u.w)}: two grid barriers and ONE
// cross-GPU all-reduce per iteration instead of three and two.  Same iterates as the classic recurrence in exact arithmetic.
template <int BS, int CW, bool SR>
__global__ void __launch_bounds__((CW + 1) * 32, 1) cg_stream(CgArgs A, StreamArgs S, P2PArgs P) {
    constexpr int ST_THREADS = (CW + 1) * 32;
    constexpr int VB = CW > 8 ? 5 : 8;  // dofs per thread and batch in the vector phases (register budget: 152 vs 224)
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ double sh[ST_THREADS / 32];
    __shared__ double sh4[4];
    __shared__ __align__(8) uint64_t full[ST_MAX_CW][ST_MAX_DEPTH], empty[ST_MAX_CW][ST_MAX_DEPTH];
    __shared__ int s_width[ST_MAX_CW][ST_MAX_DEPTH];  // blocks per row of the slice sitting in a slot
    constexpr int CDM = BS == 3 ? 6 : BS;             // most coarse dofs per aggregate (translations + rotations in 3D)
    __shared__ double s_wp[ST_MAX_CW + 1][CDM];       // two-level preconditioner: chunk sums of w = Z^T r
    __shared__ double s_y[(ST_MAX_CW + 1) * CDM];     // single-reduction two-level solver: coarse values of this CTA's aggregates
    constexpr int BB = BS * BS;
    constexpr int PS = BS == 3 ? 4 : BS;  // stride of a node in the padded search direction
    const int tid = threadIdx.x;
    const int64_t gtid = blockIdx.x * (int64_t)ST_THREADS + tid;
    const int64_t gsz = gridDim.x * (int64_t)ST_THREADS;
    const int nb = gridDim.x;
    double* part = A.partials;
    const int ps = A.part_stride;
    const bool mg = P.n_ranks > 0;
    unsigned int repoch = 0, hepoch = 0;
    if (mg) {
        repoch = (unsigned int)P.epochs[0];
        hepoch = (unsigned int)P.epochs[1];
    }
    auto pad_of = [](int64_t i) -> int64_t { return BS == 3 ? (i / 3) * 4 + (i % 3) : i; };  // dof -> index in p_pad
    const bool profiling = A.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long tprev = 0;
    auto prof = [&](int k) {
        if (profiling) {
            const long long tn = clock64();
            A.prof[k] += tn - tprev;
            tprev = tn;
        }
    };

    // grid-wide reduction + barrier of NV CTA partials (identical in all threads of the CTA on entry)
    // (a flag-word all-to-all among the CTAs instead of grid.sync + partials was measured: 50.2 vs 46.4 us per iteration)
    unsigned int n_red = 0;
    auto grid_reduce = [&](auto& v) {
        constexpr int NV = sizeof(v) / sizeof(double);
        // two row sets, alternating: a fast CTA writing reduction n+1 must not overwrite what a slow one still reads of n
        double* pr = part + (size_t)((n_red++ & 1u) * 4) * ps;
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) pr[k * ps + blockIdx.x] = v[k];
        }
        grid.sync();
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = sum_partials<ST_THREADS>(pr + k * ps, nb, sh);
    };

    // ---- the rings
    const int w = tid >> 5, lane = tid & 31;
    const int NCW = S.n_cw, D = S.depth;
    const bool is_producer = w == CW;                  // last warp; its lane l feeds consumer warp l
    const int cw = is_producer ? lane : w;             // consumer warp this thread consumes for / produces for
    const bool active = cw < NCW && (!is_producer || lane < NCW);
    const size_t val_bytes = (size_t)S.slot_blocks * BB * C * sizeof(double);
    const size_t slot_bytes = val_bytes + (size_t)S.slot_blocks * C * sizeof(int32_t);
    // slices of consumer warp cw in this CTA: (blockIdx + k*grid) * NCW + cw, k = 0 .. n_my-1
    const long long first = (long long)blockIdx.x * NCW + cw, step = (long long)gridDim.x * NCW;
    const int n_my = active && first < S.n_slices ? (int)((S.n_slices - 1 - first) / step) + 1 : 0;
    if (tid == 0) {
        for (int a = 0; a < ST_MAX_CW; ++a)
            for (int k = 0; k < ST_MAX_DEPTH; ++k) {
                mbar_init(&full[a][k], 1);
                mbar_init(&empty[a][k], 1);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const uint64_t pol_stream = l2_policy_evict_first();
    // producer lane: fill number f of its consumer warp -> slice first + (f mod n_my) * step, slot f mod D
    int64_t nx_b0 = 0, nx_b1 = 0;  // producer lane: block range of the NEXT fill's slice, fetched one fill ahead
    auto prefetch_range = [&](long long f) {
        const long long sl = first + (f % n_my) * step;
        nx_b0 = A.slice_ptr[sl];
        nx_b1 = A.slice_ptr[sl + 1];
    };
    auto fill = [&](long long f) {
        const int slot = (int)(f % D);
        const int64_t b0 = nx_b0;
        const uint32_t nblk = (uint32_t)(nx_b1 - nx_b0);
        prefetch_range(f + 1);
        unsigned char* dst = dyn + ((size_t)cw * D + slot) * slot_bytes;
        s_width[cw][slot] = (int)nblk;
        mbar_expect_tx(&full[cw][slot], nblk * (uint32_t)(BB * C * sizeof(double) + C * sizeof(int32_t)));  // release: s_width is visible to the waiters
        if (nblk) {
            bulk_g2s(dst, A.val + b0 * BB * C, nblk * (uint32_t)(BB * C * sizeof(double)), &full[cw][slot], pol_stream);
            bulk_g2s(dst + val_bytes, A.col + b0 * C, nblk * (uint32_t)(C * sizeof(int32_t)), &full[cw][slot], pol_stream);
        }
    };
    long long f_next = 0;   // producer lane: next fill to issue
    long long q_done = 0;   // consumer warp: fills consumed so far
    if (is_producer && n_my > 0) {
        prefetch_range(0);
        for (; f_next < D; ++f_next) fill(f_next);
    }

    // ---- two-level preconditioner (precond = 2): z = D^-1 r + Z y, y = E^-1 w, w = Z^T r, for the r every CTA has
    //      just finished writing.  z is never stored: r.z = sum r^2 d + w.y, so this routine only leaves y in memory
    //      and returns the thread's share of w.y; the p update forms z_i = d_i r_i + y[agg(i)] on the fly.
    //      Two more grid barriers per application (r complete -> w; w complete -> y), the third is the reduction
    //      that follows anyway.
    const bool two_level = A.precond == 2;
    const int CD = A.co.cd;
    const bool rbm = BS == 3 && CD == 6;
    long long tca = 0;
    auto cprof = [&](int k) {  // aux counters 8..11 (cycles of block 0): barrier, w = Z^T r (stand-alone pass), barrier, y = E^-1 w
        if (profiling) {
            const long long tn = clock64();
            A.prof[k] += tn - tca;
            tca = tn;
        }
    };
    // aggregates are dealt over ALL CTAs: AG_PER_CTA per CTA and pass, SPLIT warps share one aggregate (512 aggregates on
    // 148 CTAs: 4 x 3 warps; 256 aggregates: 2 x 6 warps).  Fixed by (n_agg, grid) alone, so the summation order is too.
    const int AG_PER_CTA = two_level ? max(1, min(CW + 1, (A.co.n_agg + (int)gridDim.x - 1) / (int)gridDim.x)) : 1;
    const int SPLIT = (CW + 1) / AG_PER_CTA;
    // adds node nd's share of w = Z^T r to the lane's accumulators: [sum r | sum rho x r]
    auto w_accumulate = [&](double (&acc)[CDM], int64_t nd, const double (&rv)[BS]) {
#pragma unroll
        for (int c = 0; c < BS; ++c) acc[c] += rv[c];
        if constexpr (BS == 3) {
            if (rbm) {
                const double* rp = A.co.rho + nd * 3;
                const double r0 = rp[0], r1 = rp[1], r2 = rp[2];
                acc[3] += r1 * rv[2] - r2 * rv[1];
                acc[4] += r2 * rv[0] - r0 * rv[2];
                acc[5] += r0 * rv[1] - r1 * rv[0];
            }
        }
    };
    // the SPLIT chunk sums of an aggregate meet in shared memory and are added in chunk order (both barriers inside)
    auto w_combine = [&](const double (&acc)[CDM], bool active, int base_a) {
        if (active) {
#pragma unroll
            for (int c = 0; c < CDM; ++c) {
                double v = acc[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) s_wp[w][c] = v;
            }
        }
        __syncthreads();
        if (tid < AG_PER_CTA * CD && base_a + tid / CD < A.co.n_agg) {
            const int m2 = tid / CD, c = tid % CD;
            double v = 0.0;
            for (int k2 = 0; k2 < SPLIT; ++k2) v += s_wp[m2 * SPLIT + k2][c];
            A.co.w[(base_a + m2) * CD + c] = v;
        }
        __syncthreads();
    };
    // stand-alone w = Z^T r for the r every CTA has just finished writing (prologue): barrier, then the gather pass.
    // SPLIT warps of one CTA share an aggregate (enough warps in flight to hide the dependent load levels).
    auto coarse_w_pass = [&]() {
        const CoarseArgs& G = A.co;
        if (profiling) tca = clock64();
        grid.sync();
        cprof(8);
        for (int base_a = (int)blockIdx.x * AG_PER_CTA; base_a < G.n_agg; base_a += (int)gridDim.x * AG_PER_CTA) {
            const int m = w / SPLIT, ch = w % SPLIT, a = base_a + m;
            const bool active = m < AG_PER_CTA && a < G.n_agg;
            double acc[CDM];
#pragma unroll
            for (int c = 0; c < CDM; ++c) acc[c] = 0.0;
            if (active) {
                const int q0 = G.agg_ptr[a], q1 = G.agg_ptr[a + 1];
                const int len = (q1 - q0 + SPLIT - 1) / SPLIT;
                const int b0 = q0 + ch * len, b1 = b0 + len < q1 ? b0 + len : q1;
                for (int q = b0 + lane; q < b1; q += 128) {  // four nodes per lane and trip: ids first, then their residuals
                    int64_t nd[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) nd[u] = q + 32 * u < b1 ? G.agg_nodes[q + 32 * u] : -1;
                    double rv[4][BS];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int c = 0; c < BS; ++c) rv[u][c] = nd[u] >= 0 ? A.r[nd[u] * BS + c] : 0.0;
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (nd[u] >= 0) w_accumulate(acc, nd[u], rv[u]);
                }
            }
            w_combine(acc, active, base_a);
        }
        cprof(9);
    };
    // y = E^-1 w once w is complete: barrier, then one warp per row of the dense inverse (L2-resident; 16-byte loads,
    // 12 + 12 of them in flight per lane).  z is never stored: r.z = sum r^2 d + w.y, so this leaves y in memory and
    // returns the thread's share of w.y; the p update forms z_i = d_i r_i + (Z y)_i on the fly.
    unsigned int co_apps = 0;  // applications of the global coarse level in this launch (its epochs continue across launches)
    // row k of y = E^-1 w by one warp (the dense inverse is L2-resident; 16-byte loads, 12 + 12 of them in flight per lane);
    // the result is valid in every lane.  w (and y in the p update) are ordinary loads: the grid barrier in front of them makes
    // the other CTAs' stores visible, and the 12 KB vector then comes from L1 for all but the first warp of an SM (as L2-only
    // loads every row's warp fetched its own copy: as many L2 bytes again as the dense inverse itself)
    auto einv_row_dot = [&](int k) -> double {
        const CoarseArgs& G = A.co;
        double acc = 0.0;
        if ((G.nc & 1) == 0) {  // rows are 16-byte aligned
            const double2* row = reinterpret_cast<const double2*>(G.Einv + (size_t)k * G.nc);
            const double2* wv = reinterpret_cast<const double2*>(G.w);
            const int n2 = G.nc >> 1;
            int j = lane;
            for (; j + 11 * 32 < n2; j += 12 * 32) {
                double2 e[12], ww[12];
#pragma unroll
                for (int u = 0; u < 12; ++u) {
                    e[u] = __ldg(row + j + 32 * u);
                    ww[u] = wv[j + 32 * u];
                }
#pragma unroll
                for (int u = 0; u < 12; ++u) acc += e[u].x * ww[u].x + e[u].y * ww[u].y;
            }
            for (; j < n2; j += 32) {
                const double2 e = __ldg(row + j), ww = wv[j];
                acc += e.x * ww.x + e.y * ww.y;
            }
        } else {
            const double* row = G.Einv + (size_t)k * G.nc;
            for (int j = lane; j < G.nc; j += 32) acc += __ldg(row + j) * G.w[j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        return acc;
    };
    // y = E^-1 w once w is complete: barrier, then one warp per row of the dense inverse.  z is never stored:
    // r.z = sum r^2 d + w.y, so this leaves y in memory and returns the thread's share of w.y; the p update forms
    // z_i = d_i r_i + (Z y)_i on the fly.
    auto coarse_solve = [&]() -> double {
        const CoarseArgs& G = A.co;
        const int gw = (int)(gtid >> 5), nw = (int)(gsz >> 5);
        if (profiling) tca = clock64();
        grid.sync();
        cprof(10);
        unsigned int e2 = 0;
        if (G.glob) {
            // w2 = T^T w of this rank's level-2 aggregates -> every rank (LL words); one CTA turns the gathered vector into
            // plain doubles for the row dots below
            e2 = (unsigned int)(*G.w2_epoch) + co_apps + 1u;
            const int par = (int)(e2 & 1u);
            if (blockIdx.x == 0) {
                for (int t = tid; t < G.n2_own * 6; t += ST_THREADS) {
                    const int P2 = t / 6, q = t % 6;
                    double v = 0.0;
                    for (int a = G.child_ptr[P2]; a < G.child_ptr[P2 + 1]; ++a) {
                        const double* wa = G.w + (size_t)a * 6;
                        v += __ldcg(wa + q);
                        if (q >= 3) {  // m2 = m1 + d x f1
                            const double* d = G.dvec + (size_t)a * 3;
                            const double f0 = __ldcg(wa), f1 = __ldcg(wa + 1), f2 = __ldcg(wa + 2);
                            v += q == 3 ? d[1] * f2 - d[2] * f1 : q == 4 ? d[2] * f0 - d[0] * f2 : d[0] * f1 - d[1] * f0;
                        }
                    }
                    const size_t slot = (size_t)par * G.nc2 + (size_t)(G.agg2_first + P2) * 6 + q;
                    for (int r = 0; r < G.n_ranks; ++r) ll_store(G.peer_w2[r] + 2 * slot, v, e2);
                }
            }
            if (blockIdx.x == gridDim.x - 1) {
                for (int t = tid; t < G.nc2; t += ST_THREADS) G.w2_plain[t] = ll_load(G.w2_ll + 2 * ((size_t)par * G.nc2 + t), e2, A.err);
                __threadfence();
                __syncthreads();
                if (tid == 0) *(volatile unsigned int*)G.w2_flag = e2;
            }
        }
        double wy = 0.0;
        bool have_w2 = false;
        for (int k = gw; k < G.nc; k += nw) {
            double acc = einv_row_dot(k);
            if (G.glob) {
                if (!have_w2) {  // the gathered w2 is ready once the flag carries this application's epoch
                    const long long t0 = clock64();
                    while (*(volatile unsigned int*)G.w2_flag != e2) {
                        if (clock64() - t0 > 4000000000LL) {
                            *A.err = 2;
                            break;
                        }
                    }
                    have_w2 = true;
                }
                const double* grow = G.G + (size_t)k * G.nc2;
                double g = 0.0;
                for (int j = lane; j < G.nc2; j += 32) g += __ldg(grow + j) * __ldcg(G.w2_plain + j);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
                acc += g;
            }
            if (lane == 0) {
                G.y[k] = acc;
                wy += __ldcg(G.w + k) * acc;
            }
        }
        if (G.glob) ++co_apps;
        cprof(11);
        return wy;
    };

    // ---- prologue: residual, Jacobi diagonal, norms; multi-GPU: first halo push of the preconditioned residual
    double g4[4];
    cg_prologue_body<BS>(A, gtid, gsz, g4);
    if (mg) ++hepoch;
    if (two_level) {
        coarse_w_pass();
        g4[P_RZ] += coarse_solve();  // r.z = sum r^2 d (already there) + w.y; z itself is pushed / formed in the p update
    } else if (mg) {
        for (int64_t i = gtid; i < A.n; i += gsz)
            if (A.mask[i] & 2) p2p_push(P, i, A.r[i] * A.dinv[i], hepoch);
    }
    for (int64_t i = gtid, np = ((A.n + (mg ? P.n_halo_dofs : 0)) / BS) * PS; i < np; i += gsz) A.p_pad[i] = 0.0;
    if (mg)
        for (int64_t i = A.n + gtid; i < A.n + P.n_halo_dofs; i += gsz) A.x[i] = 0.0;  // the solution on the halo dofs (see the update)
#pragma unroll
    for (int k = 0; k < 4; ++k) g4[k] = block_sum<ST_THREADS>(g4[k], sh);
    grid_reduce(g4);
    if (mg) p2p_allreduce<4>(P, g4, repoch, sh4);
    const double rr0 = g4[P_RR], ff = g4[P_FF], uu = g4[P_UU];
    double rho = g4[P_RZ];
    double res = sqrt(rr0);
    const double tol = fmax(A.reltol * res, A.abstol);
    double rho_prev = 1.0;
    long long it = 0;
    if (profiling) tprev = clock64();
    const int lrow = lane & 7, qpart = lane >> 3;

    // ---- Ap = K v from the shared-memory rings (v = the padded vector p_pad), returns the thread's share of v.Ap
    auto spmv_phase = [&]() -> double {
        double d = 0.0;
        if (is_producer) {
            // stay DEPTH fills ahead: n_my fills per SpMV phase, each as soon as its slot has been released
            if (n_my > 0)
                for (int k = 0; k < n_my; ++k, ++f_next) {
                    mbar_wait(&empty[cw][f_next % D], (uint32_t)((f_next / D - 1) & 1), A.err);
                    fill(f_next);
                }
            // the ring now holds (or is receiving) the first D slices of the NEXT phase; the PF slices behind them go to L2
            const int pf = min(S.l2_prefetch, n_my - D);
            for (int j0 = 0; j0 < pf; j0 += 4) {
                int64_t b0[4], b1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long sl = first + ((f_next + j0 + (j0 + u < pf ? u : 0)) % n_my) * step;
                    b0[u] = A.slice_ptr[sl];
                    b1[u] = A.slice_ptr[sl + 1];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j0 + u < pf && b1[u] > b0[u]) {
                        const uint32_t nblk = (uint32_t)(b1[u] - b0[u]);
                        // same L2 class as the stream itself (evict_first)
                        bulk_prefetch_l2(A.val + b0[u] * BB * C, nblk * (uint32_t)(BB * C * sizeof(double)), pol_stream);
                        bulk_prefetch_l2(A.col + b0[u] * C, nblk * (uint32_t)(C * sizeof(int32_t)), pol_stream);
                    }
            }
        } else if (n_my > 0) {
            const long long tc0 = A.prof != nullptr ? clock64() : 0;
            for (int k = 0; k < n_my; ++k, ++q_done) {
                const int slot = (int)(q_done % D);
                mbar_wait(&full[cw][slot], (uint32_t)((q_done / D) & 1), A.err);
                const int width = s_width[cw][slot];
                const unsigned char* stg = dyn + ((size_t)cw * D + slot) * slot_bytes;
                const double* sv = reinterpret_cast<const double*>(stg) + lrow;
                const int32_t* sc = reinterpret_cast<const int32_t*>(stg + val_bytes) + lrow;
                const int64_t row = (first + k * step) * C + lrow;
                const bool own = qpart == 0 && row < A.n_rows;
                double po[BS];  // the row's own v (for v.Ap), fetched in the same round trip as the gathers
#pragma unroll
                for (int r = 0; r < BS; ++r) po[r] = 0.0;
                if (own) ld_node<BS>(A.p_pad + row * PS, po);
                double acc[BS];
#pragma unroll
                for (int r = 0; r < BS; ++r) acc[r] = 0.0;
#pragma unroll 4
                for (int blk = qpart; blk < width; blk += 4) {
                    const int64_t cn = sc[blk * C];
                    double pv[BS];
                    ld_node<BS>(A.p_pad + cn * PS, pv);
#pragma unroll
                    for (int r = 0; r < BS; ++r)
#pragma unroll
                        for (int q = 0; q < BS; ++q) acc[r] += sv[(blk * BB + r * BS + q) * C] * pv[q];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[cw][slot]);  // every lane has read its part of the slot: release it
#pragma unroll
                for (int r = 0; r < BS; ++r) {
                    acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 8);
                    acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], 16);
                }
                if (own) {  // rows of fixed dofs are not masked here: v is zero there and the update phase skips them
#pragma unroll
                    for (int r = 0; r < BS; ++r) {
                        A.Ap[row * BS + r] = acc[r];
                        d += po[r] * acc[r];
                    }
                }
            }
            if (A.prof != nullptr && lane == 0) A.prof[16 + blockIdx.x * CW + w] += clock64() - tc0;  // per-warp SpMV cycles (diagnostics)
        }
        return d;
    };

    // fused residual update (two-level): this warp's chunk of its aggregate, fixed for the whole solve
    const int fw_base_a = (int)blockIdx.x * AG_PER_CTA;
    const bool fw_cta = two_level && fw_base_a < A.co.n_agg;  // AG_PER_CTA = ceil(n_agg / grid): a single pass covers all aggregates
    bool fw_active = false;
    int fw_b0 = 0, fw_b1 = 0;
    int fw_nid[4] = {-1, -1, -1, -1};
    if (fw_cta && (A.co.fused != 0 || SR)) {
        const int m = w / SPLIT, ch = w % SPLIT, a = fw_base_a + m;
        fw_active = m < AG_PER_CTA && a < A.co.n_agg;
        if (fw_active) {
            const int q0 = A.co.agg_ptr[a], q1 = A.co.agg_ptr[a + 1];
            const int len = (q1 - q0 + SPLIT - 1) / SPLIT;
            fw_b0 = q0 + ch * len;
            fw_b1 = fw_b0 + len < q1 ? fw_b0 + len : q1;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = fw_b0 + lane + 32 * u;
                fw_nid[u] = q < fw_b1 ? A.co.agg_nodes[q] : -1;
            }
        }
    }
    // classic two-level solver in 3D: the fused update runs one lane per DOF of the chunk's node list (dof t = 32 j + lane of the
    // chunk <-> node t / 3, component t % 3 = (lane + 2 j) % 3): the list is ascending, so consecutive lanes touch consecutive
    // doubles inside every run of consecutive nodes -- one lane per node (24-byte stride, three instructions per vector over the
    // same lines) needed about three times the L1 wavefronts.  Nodes of the lane's dofs in the chunk's first 128 nodes:
    int fw_dn[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) fw_dn[j] = -1;
    if constexpr (BS == 3 && !SR) {
        if (fw_active) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const int q = fw_b0 + (32 * j + lane) / 3;
                fw_dn[j] = q < fw_b1 ? A.co.agg_nodes[q] : -1;
            }
        }
    }
    if constexpr (SR) {
        // ---- single-reduction CG (precond 0 / 1)
        double* __restrict__ xv = A.x;
        double* __restrict__ rv = A.r;
        double* __restrict__ pvec = A.p;
        double* __restrict__ sv2 = A.s;
        double* __restrict__ uv = A.p_pad;
        const double* __restrict__ wv = A.Ap;
        const double* __restrict__ dv = A.dinv;
        // (Z y)_c of a node of an aggregate whose coarse values are ya[0 .. CD): translation + omega x rho
        auto zy_node = [&](const double* ya, int64_t nd, double (&zc)[BS]) {
#pragma unroll
            for (int c = 0; c < BS; ++c) zc[c] = ya[c];
            if constexpr (BS == 3) {
                if (rbm) {
                    const double* rp = A.co.rho + nd * 3;
                    const double r0 = rp[0], r1 = rp[1], r2 = rp[2];
                    zc[0] += ya[4] * r2 - ya[5] * r1;
                    zc[1] += ya[5] * r0 - ya[3] * r2;
                    zc[2] += ya[3] * r1 - ya[4] * r0;
                }
            }
        };
        // u = M^-1 r of the initial residual into the padded vector (owned), pushed to the neighbours; s = 0
        if (two_level) {  // y = E^-1 Z^T r0 is in memory (prologue above); one thread per node
            for (int64_t nd = gtid; nd < A.n_rows; nd += gsz) {
                double ya[CDM], zc[BS];
                const double* yp = A.co.y + (size_t)A.co.agg[nd] * CD;
                for (int c = 0; c < CD; ++c) ya[c] = yp[c];
                zy_node(ya, nd, zc);
#pragma unroll
                for (int c = 0; c < BS; ++c) {
                    const int64_t i = nd * BS + c;
                    const double ui = dv[i] != 0.0 ? rv[i] * dv[i] + zc[c] : 0.0;
                    uv[pad_of(i)] = ui;
                    sv2[i] = 0.0;
                    if (mg && (A.mask[i] & 2)) p2p_push(P, i, ui, hepoch);
                }
            }
        } else
        for (int64_t i = gtid; i < A.n; i += gsz) {
            const double ui = rv[i] * dv[i];
            uv[pad_of(i)] = ui;
            sv2[i] = 0.0;  // (multi-GPU: the prologue above has already pushed u of the interface dofs with this epoch)
        }
        if (mg)
            for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz) {
                uv[pad_of(A.n + h)] = ll_load(P.zh + 2 * h, hepoch, P.err);
                pvec[A.n + h] = 0.0;
            }
        grid.sync();
        double dl[1] = {block_sum<ST_THREADS>(spmv_phase(), sh)};  // delta_0 = u.K u
        grid_reduce(dl);
        if (mg) p2p_allreduce<1>(P, dl, repoch, sh4);
        double gamma = rho, alpha = gamma / dl[0], beta = 0.0;
        if (profiling) tprev = clock64();
        while (!(it >= A.maxiter || res <= tol || res != res)) {
            if (*(volatile int*)A.err == 3) break;
            ++hepoch;
            double s3[3] = {0.0, 0.0, 0.0};
            if (two_level) {
                // ---- two-level: the same recurrences in AGGREGATE order (this warp's chunk of its aggregate, as in the
                //      classic solver's fused update), so that w = Z^T r of the new residual is accumulated on the way;
                //      then, once w is complete, every CTA solves for the coarse values of ITS aggregates only
                //      (y = rows of E^-1 times w) and forms u = D^-1 r + Z y for their nodes.  Three grid barriers and one
                //      cross-GPU all-reduce per iteration (classic two-level: four and two).
                double acc[CDM];
#pragma unroll
                for (int c = 0; c < CDM; ++c) acc[c] = 0.0;
                auto v1_node = [&](int64_t nd) {
                    double u[BS], pp[BS], ww[BS], ss[BS], x[BS], rr[BS], di[BS], rn[BS];
                    ld_node<BS>(uv + nd * PS, u);
#pragma unroll
                    for (int c = 0; c < BS; ++c) {
                        const int64_t i = nd * BS + c;
                        pp[c] = pvec[i];
                        ww[c] = wv[i];
                        ss[c] = sv2[i];
                        x[c] = xv[i];
                        rr[c] = rv[i];
                        di[c] = dv[i];
                    }
#pragma unroll
                    for (int c = 0; c < BS; ++c) {
                        const int64_t i = nd * BS + c;
                        const double pn = u[c] + beta * pp[c];
                        const double sn = ww[c] + beta * ss[c];
                        pvec[i] = pn;
                        sv2[i] = sn;
                        xv[i] = x[c] + alpha * pn;
                        const double ri = di[c] != 0.0 ? rr[c] - alpha * sn : 0.0;
                        rv[i] = ri;
                        rn[c] = ri;
                        s3[0] += ri * ri;
                        s3[1] += ri * (ri * di[c]);
                    }
                    w_accumulate(acc, nd, rn);
                };
                if (fw_cta) {
                    if (fw_active) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (fw_nid[k] >= 0) v1_node(fw_nid[k]);
                        for (int q = fw_b0 + 128 + lane; q < fw_b1; q += 32) v1_node(A.co.agg_nodes[q]);  // chunks longer than 128 nodes
                    }
                    w_combine(acc, fw_active, fw_base_a);
                }
                if (mg)  // halo dofs: the owner's p and x recurrences (their old u is still in the padded vector)
                    for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz) {
                        const int64_t i = A.n + h;
                        const double pn = uv[pad_of(i)] + beta * pvec[i];
                        pvec[i] = pn;
                        xv[i] += alpha * pn;
                    }
                prof(4);
                grid.sync();  // w = Z^T r complete
                prof(5);
                if (fw_cta) {
                    for (int k = w; k < AG_PER_CTA * CD; k += CW + 1) {  // one warp per coarse row of this CTA's aggregates
                        const int a = fw_base_a + k / CD;
                        if (a < A.co.n_agg) {
                            const int gk = a * CD + k % CD;
                            const double yk = einv_row_dot(gk);
                            if (lane == 0) {
                                s_y[k] = yk;
                                s3[1] += __ldcg(A.co.w + gk) * yk;  // r.u = sum r^2 d + w.y
                            }
                        }
                    }
                    __syncthreads();
                    if (fw_active) {
                        double ya[CDM];
                        for (int c = 0; c < CD; ++c) ya[c] = s_y[(w / SPLIT) * CD + c];
                        auto v2_node = [&](int64_t nd) {
                            double zc[BS];
                            zy_node(ya, nd, zc);
#pragma unroll
                            for (int c = 0; c < BS; ++c) {
                                const int64_t i = nd * BS + c;
                                const double d = dv[i];
                                const double un = d != 0.0 ? rv[i] * d + zc[c] : 0.0;
                                uv[nd * PS + c] = un;
                                if (mg && (A.mask[i] & 2)) p2p_push(P, i, un, hepoch);
                            }
                        };
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (fw_nid[k] >= 0) v2_node(fw_nid[k]);
                        for (int q = fw_b0 + 128 + lane; q < fw_b1; q += 32) v2_node(A.co.agg_nodes[q]);
                    }
                }
                if (mg)
                    for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz) uv[pad_of(A.n + h)] = ll_load(P.zh + 2 * h, hepoch, P.err);
            } else {
            constexpr int VS = 3;  // dofs per thread and batch: seven vectors are live per dof (128 registers per thread)
            for (int64_t i0 = gtid; i0 < A.n; i0 += VS * gsz) {
                double u[VS], pp[VS], ww[VS], ss[VS], x[VS], rr[VS], di[VS];
#pragma unroll
                for (int k = 0; k < VS; ++k) {  // loads of a batch first (tail lanes re-read the last dof: no branches)
                    const int64_t i = i0 + k * gsz < A.n ? i0 + k * gsz : A.n - 1;
                    u[k] = uv[pad_of(i)];
                    pp[k] = pvec[i];
                    ww[k] = wv[i];
                    ss[k] = sv2[i];
                    x[k] = xv[i];
                    rr[k] = rv[i];
                    di[k] = dv[i];
                }
#pragma unroll
                for (int k = 0; k < VS; ++k) {
                    const int64_t i = i0 + k * gsz;
                    if (i < A.n) {
                        const double pn = u[k] + beta * pp[k];
                        const double sn = ww[k] + beta * ss[k];
                        pvec[i] = pn;
                        sv2[i] = sn;
                        xv[i] = x[k] + alpha * pn;
                        const double ri = di[k] != 0.0 ? rr[k] - alpha * sn : 0.0;  // fixed dofs: d = 0 (the SpMV does not mask)
                        rv[i] = ri;
                        const double un = ri * di[k];
                        uv[pad_of(i)] = un;
                        s3[0] += ri * ri;
                        s3[1] += ri * un;
                        if (mg && (A.mask[i] & 2)) p2p_push(P, i, un, hepoch);
                    }
                }
            }
            if (mg)  // halo dofs: the same p and x recurrences as on the owner (bitwise), then the neighbours' new u
                for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz) {
                    const int64_t i = A.n + h, j = pad_of(i);
                    const double pn = uv[j] + beta * pvec[i];
                    pvec[i] = pn;
                    xv[i] += alpha * pn;
                    uv[j] = ll_load(P.zh + 2 * h, hepoch, P.err);
                }
            }
            prof(0);
            grid.sync();
            prof(1);
            s3[2] = spmv_phase();
#pragma unroll
            for (int k = 0; k < 3; ++k) s3[k] = block_sum<ST_THREADS>(s3[k], sh);
            prof(2);
            grid_reduce(s3);
            prof(3);
            if (mg) p2p_allreduce<3>(P, s3, repoch, sh4);
            beta = s3[1] / gamma;
            alpha = s3[1] / (s3[2] - beta * s3[1] / alpha);
            rho_prev = gamma;
            gamma = s3[1];
            rho = gamma;
            res = sqrt(s3[0]);
            ++it;
            prof(6);
        }
    }

    while (!SR && !(it >= A.maxiter || res <= tol || res != res)) {  // a non-finite residual (breakdown) ends the solve: err = 4
        // ---- p = z + beta p (owned dofs; halo dofs from the neighbours' pushes of epoch hepoch)
        {
            const double beta = rho / rho_prev;
            const double* __restrict__ rv = A.r;
            const double* __restrict__ dv = A.dinv;
            double* __restrict__ pv = A.p_pad;
            bool nodewise = false;
            if constexpr (BS == 3) {
                if (two_level) {
                    // one thread per NODE: its aggregate id, the aggregate's 3 (or 6) coarse values and the node's rho are
                    // fetched once for the three dofs; p travels as one 32-byte sector in and out
                    nodewise = true;
                    const int32_t* __restrict__ agv = A.co.agg;
                    constexpr int NB = 2;  // nodes per thread and trip (3 would cover the 1 M-tet cube in one trip but spills: 54 live doubles)
                    for (int64_t n0 = gtid; n0 < A.n_rows; n0 += NB * gsz) {
                        double ra[NB][3], da[NB][3], pa[NB][3], rh[NB][3];
                        int ag[NB];
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            const int64_t nd = n0 + u * gsz < A.n_rows ? n0 + u * gsz : A.n_rows - 1;
                            ag[u] = agv[nd];
                            ld_node<BS>(pv + nd * PS, pa[u]);
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) {
                                ra[u][cc] = rv[nd * 3 + cc];
                                da[u][cc] = dv[nd * 3 + cc];
                                rh[u][cc] = rbm ? A.co.rho[nd * 3 + cc] : 0.0;
                            }
                        }
                        double ty[NB][3], om[NB][3];
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            const double* ya = A.co.y + (size_t)ag[u] * CD;
#pragma unroll
                            for (int cc = 0; cc < 3; ++cc) {
                                ty[u][cc] = ya[cc];
                                om[u][cc] = rbm ? ya[3 + cc] : 0.0;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            const int64_t nd = n0 + u * gsz;
                            if (nd < A.n_rows) {
                                const double zc[3] = {ty[u][0] + om[u][1] * rh[u][2] - om[u][2] * rh[u][1],   // t + omega x rho
                                                      ty[u][1] + om[u][2] * rh[u][0] - om[u][0] * rh[u][2],
                                                      ty[u][2] + om[u][0] * rh[u][1] - om[u][1] * rh[u][0]};
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) {
                                    const double zi = da[u][cc] != 0.0 ? ra[u][cc] * da[u][cc] + zc[cc] : 0.0;
                                    pv[nd * PS + cc] = zi + beta * pa[u][cc];
                                    if (mg && (A.mask[nd * 3 + cc] & 2)) p2p_push(P, nd * 3 + cc, zi, hepoch);
                                }
                            }
                        }
                    }
                }
            }
            // one CTA per SM: batches of VB independent dofs per thread keep enough loads in flight
            for (int64_t i0 = gtid; !nodewise && i0 < A.n; i0 += VB * gsz) {
                double a[VB], b[VB], c[VB], yc[VB];
#pragma unroll
                for (int u = 0; u < VB; ++u) {  // loads of a batch first (tail lanes re-read the last dof: no branches)
                    const int64_t i = i0 + u * gsz < A.n ? i0 + u * gsz : A.n - 1;
                    a[u] = rv[i];
                    b[u] = dv[i];
                    c[u] = pv[pad_of(i)];
                    yc[u] = 0.0;
                    if (two_level) {
                        const int64_t nd = i / BS;
                        const int cc = (int)(i % BS);
                        const double* ya = A.co.y + (size_t)A.co.agg[nd] * CD;
                        yc[u] = ya[cc];
                        if constexpr (BS == 3) {
                            if (rbm) {  // (omega x rho)_c = omega_{c+1} rho_{c+2} - omega_{c+2} rho_{c+1}
                                const int c1 = cc == 2 ? 0 : cc + 1, c2 = cc == 0 ? 2 : cc - 1;
                                const double* rp = A.co.rho + nd * 3;
                                yc[u] += ya[3 + c1] * rp[c2] - ya[3 + c2] * rp[c1];
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < VB; ++u) {
                    const int64_t i = i0 + u * gsz;
                    if (i < A.n) {
                        const double zi = b[u] != 0.0 ? a[u] * b[u] + yc[u] : 0.0;  // z = D^-1 r (+ Z y), zero at fixed dofs
                        pv[pad_of(i)] = zi + beta * c[u];
                        if (two_level && mg && (A.mask[i] & 2)) p2p_push(P, i, zi, hepoch);  // Jacobi pushed z with the r update
                    }
                }
            }
            if (mg) {
                for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz) {
                    const int64_t j = pad_of(A.n + h);
                    pv[j] = ll_load(P.zh + 2 * h, hepoch, P.err) + beta * pv[j];
                }
            }
        }
        prof(0);
        grid.sync();
        prof(1);
        // ---- Ap = K p from the shared-memory rings, partial p.Ap
        const double d = spmv_phase();
        double pAp[1] = {block_sum<ST_THREADS>(d, sh)};
        prof(2);
        grid_reduce(pAp);
        prof(3);
        if (*(volatile int*)A.err == 3) break;  // a stage never arrived (set before the barrier, so every CTA sees it)
        if (mg) p2p_allreduce<1>(P, pAp, repoch, sh4);
        const double alpha = rho / pAp[0];
        // ---- x += alpha p ; r -= alpha Ap ; multi-GPU: push z of the interface dofs right away
        ++hepoch;
        double s2[2] = {0.0, 0.0};
        const bool fused_w = two_level && A.co.fused != 0;
        if (fused_w) {
            // the same update in AGGREGATE order, so that w = Z^T r of the new residual is accumulated on the way
            // (no second pass over r, no barrier between the update and the gather).  A warp keeps the same chunk of
            // the same aggregate for the whole solve: the ids of its first 128 nodes live in registers (fw_nid), so
            // an update is one round trip to the five vectors per two nodes instead of three dependent ones.
            const CoarseArgs& G = A.co;
            if (fw_cta) {
                double acc[CDM];
#pragma unroll
                for (int c = 0; c < CDM; ++c) acc[c] = 0.0;
                // (the nodes' rho travels in the same round trip as the five vectors: fetched inside the accumulation, after the
                // stores of x and r, it cost a dependent L2 round trip per node)
                auto update_pair = [&](const int64_t (&nd)[2]) {
                    double x[2][BS], pp[2][BS], rr[2][BS], ap[2][BS], di[2][BS], rh[2][BS];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int64_t n0 = nd[u] >= 0 ? nd[u] : 0;
                        ld_node<BS>(A.p_pad + n0 * PS, pp[u]);
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            const int64_t i = n0 * BS + c;
                            x[u][c] = A.x[i];
                            rr[u][c] = A.r[i];
                            ap[u][c] = A.Ap[i];
                            di[u][c] = A.dinv[i];
                            rh[u][c] = rbm ? G.rho[i] : 0.0;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (nd[u] < 0) continue;
                        double rn[BS];
#pragma unroll
                        for (int c = 0; c < BS; ++c) {
                            const int64_t i = nd[u] * BS + c;
                            A.x[i] = x[u][c] + alpha * pp[u][c];
                            const double ri = di[u][c] != 0.0 ? rr[u][c] - alpha * ap[u][c] : 0.0;
                            A.r[i] = ri;
                            rn[c] = ri;
                            s2[0] += ri * ri;
                            s2[1] += ri * (ri * di[u][c]);
                        }
#pragma unroll
                        for (int c = 0; c < BS; ++c) acc[c] += rn[c];
                        if constexpr (BS == 3) {
                            if (rbm) {  // the same expressions as w_accumulate
                                acc[3] += rh[u][1] * rn[2] - rh[u][2] * rn[1];
                                acc[4] += rh[u][2] * rn[0] - rh[u][0] * rn[2];
                                acc[5] += rh[u][0] * rn[1] - rh[u][1] * rn[0];
                            }
                        }
                    }
                };
                // three trips = 96 consecutive dofs of the chunk's list; the lane's component in trip u is (c0 + SL[u]) % 3 with
                // c0 = lane % 3, so "slot" k of the accumulators stands for component (c0 + k) % 3 until the end of the pass.
                // r_c enters (rho x r) twice: + rho_{c+2} r_c in component c + 1 and - rho_{c+1} r_c in component c + 2.
                const int c0 = lane % 3;
                double accT[3] = {0.0, 0.0, 0.0}, accR[3] = {0.0, 0.0, 0.0};
                auto update_group = [&](const int (&nd)[3]) {
                    constexpr int SL[3] = {0, 2, 1};
                    double x[3], pp[3], rr[3], ap[3], di[3], ra[3], rb[3];
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int cu = c0 + SL[u] >= 3 ? c0 + SL[u] - 3 : c0 + SL[u];
                        const int c1 = cu == 2 ? 0 : cu + 1, c2 = cu == 0 ? 2 : cu - 1;
                        const int64_t n0 = nd[u] >= 0 ? nd[u] : 0;
                        const int64_t i = n0 * 3 + cu;
                        x[u] = A.x[i];
                        rr[u] = A.r[i];
                        ap[u] = A.Ap[i];
                        di[u] = A.dinv[i];
                        pp[u] = A.p_pad[n0 * PS + cu];
                        ra[u] = rbm ? G.rho[n0 * 3 + c2] : 0.0;
                        rb[u] = rbm ? G.rho[n0 * 3 + c1] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (nd[u] < 0) continue;
                        const int cu = c0 + SL[u] >= 3 ? c0 + SL[u] - 3 : c0 + SL[u];
                        const int64_t i = (int64_t)nd[u] * 3 + cu;
                        A.x[i] = x[u] + alpha * pp[u];
                        const double ri = di[u] != 0.0 ? rr[u] - alpha * ap[u] : 0.0;
                        A.r[i] = ri;
                        s2[0] += ri * ri;
                        s2[1] += ri * (ri * di[u]);
                        accT[SL[u]] += ri;
                        accR[(SL[u] + 1) % 3] += ra[u] * ri;
                        accR[(SL[u] + 2) % 3] -= rb[u] * ri;
                    }
                };
                if (BS == 3 && fw_active) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (fw_dn[3 * g] >= 0) {  // (the first dof of a group is the lowest: nothing follows an empty group)
                            const int nd[3] = {fw_dn[3 * g], fw_dn[3 * g + 1], fw_dn[3 * g + 2]};
                            update_group(nd);
                        }
                    }
                    for (int t0 = 384; t0 < 3 * (fw_b1 - fw_b0); t0 += 96) {  // chunks longer than 128 nodes: ids from memory
                        int nd[3];
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int q = fw_b0 + (t0 + 32 * u + lane) / 3;
                            nd[u] = q < fw_b1 ? G.agg_nodes[q] : -1;
                        }
                        update_group(nd);
                    }
                    // slots back to components
                    acc[0] = c0 == 0 ? accT[0] : c0 == 1 ? accT[2] : accT[1];
                    acc[1] = c0 == 0 ? accT[1] : c0 == 1 ? accT[0] : accT[2];
                    acc[2] = c0 == 0 ? accT[2] : c0 == 1 ? accT[1] : accT[0];
                    if constexpr (CDM == 6) {
                        acc[3] = c0 == 0 ? accR[0] : c0 == 1 ? accR[2] : accR[1];
                        acc[4] = c0 == 0 ? accR[1] : c0 == 1 ? accR[0] : accR[2];
                        acc[5] = c0 == 0 ? accR[2] : c0 == 1 ? accR[1] : accR[0];
                    }
                } else if (fw_active) {
                    {
                        const int64_t nd[2] = {fw_nid[0], fw_nid[1]};
                        update_pair(nd);
                    }
                    if (fw_nid[2] >= 0) {
                        const int64_t nd[2] = {fw_nid[2], fw_nid[3]};
                        update_pair(nd);
                    }
                    for (int q = fw_b0 + 128 + lane; q < fw_b1; q += 64) {  // chunks longer than 128 nodes: ids from memory
                        const int64_t nd[2] = {G.agg_nodes[q], q + 32 < fw_b1 ? G.agg_nodes[q + 32] : -1};
                        update_pair(nd);
                    }
                }
                w_combine(acc, fw_active, fw_base_a);
            }
        } else
        {
            double* __restrict__ xv = A.x;
            double* __restrict__ rv = A.r;
            const double* __restrict__ pv = A.p_pad;
            const double* __restrict__ av = A.Ap;
            const double* __restrict__ dv = A.dinv;
            for (int64_t i0 = gtid; i0 < A.n; i0 += VB * gsz) {
                double x[VB], pp[VB], rr[VB], ap[VB], di[VB];
#pragma unroll
                for (int u = 0; u < VB; ++u) {
                    const int64_t i = i0 + u * gsz < A.n ? i0 + u * gsz : A.n - 1;
                    x[u] = xv[i];
                    pp[u] = pv[pad_of(i)];
                    rr[u] = rv[i];
                    ap[u] = av[i];
                    di[u] = dv[i];
                }
#pragma unroll
                for (int u = 0; u < VB; ++u) {
                    const int64_t i = i0 + u * gsz;
                    if (i < A.n) {
                        xv[i] = x[u] + alpha * pp[u];
                        const double ri = di[u] != 0.0 ? rr[u] - alpha * ap[u] : 0.0;  // fixed dofs: d = 0 (the SpMV does not mask)
                        rv[i] = ri;
                        const double zi = ri * di[u];
                        s2[0] += ri * ri;
                        s2[1] += ri * zi;
                        if (mg && !two_level && (A.mask[i] & 2)) p2p_push(P, i, zi, hepoch);
                    }
                }
            }
        }
        // multi-GPU: x on the halo dofs is accumulated from the halo search directions with the same alpha, in the same
        // order and with the same fused multiply-add as on the owner, so after the epilogue U is consistent across ranks
        // without any exchange (the next assembly reads the halo part of U directly)
        if (mg)
            for (int64_t h = gtid; h < P.n_halo_dofs; h += gsz) A.x[A.n + h] += alpha * A.p_pad[pad_of(A.n + h)];
        if (two_level) {  // r.z = sum r^2 d + w.y
            if (!fused_w) coarse_w_pass();
            s2[1] += coarse_solve();
        }
        s2[0] = block_sum<ST_THREADS>(s2[0], sh);
        s2[1] = block_sum<ST_THREADS>(s2[1], sh);
        prof(4);
        grid_reduce(s2);
        prof(5);
        if (mg) p2p_allreduce<2>(P, s2, repoch, sh4);
        rho_prev = rho;
        rho = s2[1];
        res = sqrt(s2[0]);
        ++it;
        prof(6);
    }
    // drain the copies that are still in flight before the shared memory goes away
    if (!is_producer && n_my > 0)
        for (int k = 0; k < D; ++k, ++q_done) mbar_wait(&full[cw][q_done % D], (uint32_t)((q_done / D) & 1), A.err);

    if (mg && A.update_U)  // halo dofs: the same update as on their owner (the mask carries the free bit of halo dofs too)
        for (int64_t i = A.n + gtid; i < A.n + P.n_halo_dofs; i += gsz) {
            if (A.update_U == 1) A.U[i] += A.x[i];
            else if (A.mask[i] & 1) A.U[i] = A.x[i];
        }
    double dd1[1] = {block_sum<ST_THREADS>(cg_epilogue_body(A, gtid, gsz), sh)};
    grid_reduce(dd1);
    if (mg) p2p_allreduce<1>(P, dd1, repoch, sh4);
    if (blockIdx.x == 0 && tid == 0) {
        CgState* st = A.st;
        st->rho = rho;
        st->rho_prev = rho_prev;
        st->res = res;
        st->tol = tol;
        st->pAp = 0.0;
        st->rr0 = rr0;
        st->ff = ff;
        st->uu = uu;
        st->dd = dd1[0];
        st->it = it;
        st->done = 1;
        if (res != res) *A.err = 4;
        if (mg) {
            P.epochs[0] = repoch;
            P.epochs[1] = hepoch;
        }
        if (two_level && A.co.glob) *A.co.w2_epoch += co_apps;
    }
}

// ---- multi-launch driver kernels (single- or multi-GPU).  Sums that cross GPUs live in
// red[] (device doubles) so an in-stream ncclAllReduce can sit between a *_reduce kernel and
// its consumer.  Every kernel is a no-op once st->done is set, so the host may enqueue
// iterations ahead of the convergence check without changing the result.
template <int BS>
__global__ void __launch_bounds__(CG_THREADS) k_cg_prologue(CgArgs A) {
    __shared__ double sh[CG_THREADS / 32];
    const int64_t gtid = blockIdx.x * (int64_t)CG_THREADS + threadIdx.x;
    const int64_t gsz = gridDim.x * (int64_t)CG_THREADS;
    double s4[4];
    cg_prologue_body<BS>(A, gtid, gsz, s4);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double b = block_sum<CG_THREADS>(s4[k], sh);
        if (threadIdx.x == 0) A.partials[k * A.part_stride + blockIdx.x] = b;
    }
}

// single CTA: red[k] = sum of partial array first+k, k < count (fixed order)
__global__ void __launch_bounds__(CG_THREADS) k_reduce_partials(const double* partials, int part_stride, int nblk,
                                                               int first, int count, double* red, const CgState* st,
                                                               int gate) {
    __shared__ double sh[CG_THREADS / 32];
    if (gate && st->done) return;
    for (int k = 0; k < count; ++k) {
        const double v = sum_partials<CG_THREADS>(partials + (first + k) * part_stride, nblk, sh);
        if (threadIdx.x == 0) red[first + k] = v;
    }
}

// after the (all-)reduction of the prologue sums
__global__ void k_cg_init_state(CgArgs A, const double* red) {
    CgState* st = A.st;
    st->rr0 = red[P_RR];
    st->ff = red[P_FF];
    st->uu = red[P_UU];
    st->rho = red[P_RZ];
    st->rho_prev = 1.0;
    st->res = sqrt(red[P_RR]);
    st->tol = fmax(A.reltol * st->res, A.abstol);
    st->it = 0;
    st->dd = 0.0;
    st->done = (0 >= A.maxiter || st->res <= st->tol || st->res != st->res) ? 1 : 0;
    if (st->res != st->res) *A.err = 4;
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_update_p(CgArgs A) {
    if (A.st->done) return;
    const double beta = A.st->rho / A.st->rho_prev;
    cg_update_p_body(A, blockIdx.x * (int64_t)CG_THREADS + threadIdx.x, gridDim.x * (int64_t)CG_THREADS, beta);
}

template <int BS>
__global__ void __launch_bounds__(CG_THREADS) k_spmv_dot(CgArgs A, int gate) {
    __shared__ double sh[CG_THREADS / 32];
    if (gate && A.st->done) return;
    const int64_t gtid = blockIdx.x * (int64_t)CG_THREADS + threadIdx.x;
    const int64_t gsz = gridDim.x * (int64_t)CG_THREADS;
    double d = 0.0;
    for (int64_t t = gtid, nt = spmv_items<BS>(A); t < nt; t += gsz) d += spmv_item<BS>(A, t);
    d = block_sum<CG_THREADS>(d, sh);
    if (threadIdx.x == 0) A.partials[P_PAP * A.part_stride + blockIdx.x] = d;
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_update_xr(CgArgs A, const double* red) {
    __shared__ double sh[CG_THREADS / 32];
    if (A.st->done) return;
    const double alpha = A.st->rho / red[P_PAP];
    double s2[2];
    cg_update_xr_body(A, blockIdx.x * (int64_t)CG_THREADS + threadIdx.x, gridDim.x * (int64_t)CG_THREADS, alpha, s2);
    const double b0 = block_sum<CG_THREADS>(s2[0], sh);
    const double b1 = block_sum<CG_THREADS>(s2[1], sh);
    if (threadIdx.x == 0) {
        A.partials[P_RR * A.part_stride + blockIdx.x] = b0;
        A.partials[P_RZ * A.part_stride + blockIdx.x] = b1;
    }
}

// after the (all-)reduction of rr, rz
__global__ void k_cg_advance(CgArgs A, const double* red) {
    CgState* st = A.st;
    if (st->done) return;
    st->rho_prev = st->rho;
    st->rho = red[P_RZ];
    st->res = sqrt(red[P_RR]);
    st->it += 1;
    if (st->it >= A.maxiter || st->res <= st->tol || st->res != st->res) st->done = 1;
    if (st->res != st->res) *A.err = 4;
}

__global__ void __launch_bounds__(CG_THREADS) k_cg_epilogue(CgArgs A) {
    __shared__ double sh[CG_THREADS / 32];
    double dd = cg_epilogue_body(A, blockIdx.x * (int64_t)CG_THREADS + threadIdx.x, gridDim.x * (int64_t)CG_THREADS);
    dd = block_sum<CG_THREADS>(dd, sh);
    if (threadIdx.x == 0) A.partials[P_DD * A.part_stride + blockIdx.x] = dd;
}

__global__ void k_cg_finish(CgArgs A, const double* red) { A.st->dd = red[P_DD]; }

// ------------------------------------------------------------------------------------------------
// external loads (the step before each Newton loop: apply!, StructuralAnalyses.jl:228-241)
// ------------------------------------------------------------------------------------------------
// Unit nodal load vector of one boundary condition on triangular faces, built once on the device.
//   kind 0  GlobalLoad (GlobalLoadBoundaryConditions.jl:50-68): v * A / 3 on each face node, v = values[0..2]
//   kind 1  Pressure   (LocalLoadBoundaryConditions.jl:36-56):  -n * A / 3 * values[0], n A = 1/2 (x2-x1) x (x3-x1)
//           (TriangularFaces.jl:44-65)
// One thread per node walks the node's incident (face, corner) list in ascending face order and writes the node's
// entries once: the duplicate summation of StructuralBoundaryConditions.jl:195-220 with a fixed order, no atomics.
__global__ void k_face_load(const double* __restrict__ X, const int32_t* __restrict__ tri, const int64_t* __restrict__ nf_ptr,
                            const int32_t* __restrict__ nf_face, int64_t n_nodes, int kind, double v0, double v1, double v2,
                            double* __restrict__ F) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
    for (int64_t q = nf_ptr[i]; q < nf_ptr[i + 1]; ++q) {
        const int64_t f = nf_face[q];
        const double* a = X + 3 * (int64_t)tri[3 * f];
        const double* b = X + 3 * (int64_t)tri[3 * f + 1];
        const double* c = X + 3 * (int64_t)tri[3 * f + 2];
        const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
        const double av[3] = {0.5 * (e1[1] * e2[2] - e1[2] * e2[1]), 0.5 * (e1[2] * e2[0] - e1[0] * e2[2]),
                              0.5 * (e1[0] * e2[1] - e1[1] * e2[0])};
        if (kind == 0) {
            const double A3 = sqrt(av[0] * av[0] + av[1] * av[1] + av[2] * av[2]) / 3.0;
            f0 += v0 * A3;
            f1 += v1 * A3;
            f2 += v2 * A3;
        } else {
            f0 += -v0 * av[0] / 3.0;
            f1 += -v0 * av[1] / 3.0;
            f2 += -v0 * av[2] / 3.0;
        }
    }
    F[3 * i] = f0;
    F[3 * i + 1] = f1;
    F[3 * i + 2] = f2;
}

// F_ext = sum_k factor[k] * pattern_k, patterns stored back to back (n each), summed in pattern order
__global__ void k_combine_loads(const double* __restrict__ patterns, const double* __restrict__ factors, int n_patterns, int64_t n,
                                double* __restrict__ Fext) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double f = 0.0;
    for (int k = 0; k < n_patterns; ++k) f += factors[k] * patterns[(int64_t)k * n + i];
    Fext[i] = f;
}

// ---- halo exchange helpers (multi-GPU): pack owned values the neighbours need
__global__ void k_pack(const double* __restrict__ v, const int32_t* __restrict__ send_nodes, int64_t n_send, int bs,
                       double* __restrict__ buf, const CgState* st, int gate) {
    if (gate && st->done) return;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_send * bs) return;
    buf[i] = v[(int64_t)send_nodes[i / bs] * bs + (i % bs)];
}

__global__ void k_fill(double* v, int64_t n, double a) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = a;
}

}  // namespace onsas
