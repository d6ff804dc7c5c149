// hostsim.cpp -- TEST-ONLY CPU walk-through of the device algorithms in onsas.jl_b200/csrc.
//
// This container has no GPU, so the arithmetic (element_math.cuh, compiled here with g++) and
// the host-built tables (tables.cpp) are exercised by a serial transliteration of the kernels'
// control flow: phase A / phase B of k_assemble, spmv_row and the PCG phase order.  It is NOT
// part of the product, is never loaded by onsas.jl_b200, and is not a CPU fallback: the library
// has none.  tests/test_hostsim.py compares its output with the oracle.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../onsas.jl_b200/csrc/element_math.cuh"
#include "../../onsas.jl_b200/csrc/tables.hpp"

using namespace onsas;

namespace {
constexpr int C = SLICE_ROWS;

template <int K>
void tet_pair_k(const double X[4][3], const double U[4][3], double p0, double p1, int a, double* rec, double* out16) {
    TetCommon c;
    tet_common<K>(X, U, p0, p1, c);
    if (a == 0 && out16) tet_stress_out<K>(c, p0, p1, out16);
    tet_row<K>(c, U, a, rec, rec + 36);
}

void tet_pair(int kind, const double X[4][3], const double U[4][3], double p0, double p1, int a, double* rec, double* out16) {
    if (kind == MAT_SVK) tet_pair_k<MAT_SVK>(X, U, p0, p1, a, rec, out16);
    else if (kind == MAT_NEOHOOKEAN) tet_pair_k<MAT_NEOHOOKEAN>(X, U, p0, p1, a, rec, out16);
    else tet_pair_k<MAT_ISOLINEAR>(X, U, p0, p1, a, rec, out16);
}

template <int DIM>
void truss_pair_d(int strain_model, const double X[2][3], const double U[2][3], double Emod, double A, int a, double* rec,
                  double* se) {
    double blk[2][9], f[3];
    truss_row<DIM>(strain_model, X, U, Emod, A, a, blk, f, se);
    for (int b = 0; b < 2; ++b)
        for (int k = 0; k < DIM * DIM; ++k) rec[b * DIM * DIM + k] = blk[b][k];
    for (int r = 0; r < DIM; ++r) rec[2 * DIM * DIM + r] = f[r];
}
}  // namespace

extern "C" {

// un-assembled tets through tet_row (mirrors k_eval_tets)
int hs_eval_tets(int64_t n, const int32_t* conn, const int32_t* mat_id, const int32_t* kind, const double* params,
                 const double* xyz, const double* Uv, double* f, double* K, double* sig, double* eps) {
    for (int64_t e = 0; e < n; ++e) {
        double X[4][3], U[4][3];
        for (int k = 0; k < 4; ++k)
            for (int c = 0; c < 3; ++c) {
                X[k][c] = xyz[3 * (int64_t)conn[4 * e + k] + c];
                U[k][c] = Uv[3 * (int64_t)conn[4 * e + k] + c];
            }
        int m = mat_id ? mat_id[e] : 0;
        double out[16];
        for (int a = 0; a < 4; ++a) {
            double rec[TET_REC];
            tet_pair(kind[m], X, U, params[2 * m], params[2 * m + 1], a, rec, out);
            for (int r = 0; r < 3; ++r) {
                f[12 * e + 3 * a + r] = rec[36 + r];
                for (int b = 0; b < 4; ++b)
                    for (int q = 0; q < 3; ++q) K[144 * e + (3 * a + r) + 12 * (3 * b + q)] = rec[9 * b + 3 * r + q];
            }
        }
        const int VI[6] = {0, 1, 2, 1, 0, 0}, VJ[6] = {0, 1, 2, 2, 2, 1};
        for (int k = 0; k < 9; ++k) sig[9 * e + k] = out[k];
        for (int v = 0; v < 6; ++v) {
            eps[9 * e + VI[v] + 3 * VJ[v]] = out[9 + v];
            eps[9 * e + VJ[v] + 3 * VI[v]] = out[9 + v];
        }
    }
    return 0;
}

int hs_eval_trusses(int64_t n, int dim, int strain_model, const int32_t* conn, const int32_t* mat_id, const int32_t* kind,
                    const double* params, const double* area, const double* xyz, const double* Uv, double* f, double* K,
                    double* sig, double* eps) {
    const int N = 2 * dim;
    for (int64_t e = 0; e < n; ++e) {
        double X[2][3] = {{0}}, U[2][3] = {{0}};
        for (int k = 0; k < 2; ++k)
            for (int c = 0; c < dim; ++c) {
                X[k][c] = xyz[dim * (int64_t)conn[2 * e + k] + c];
                U[k][c] = Uv[dim * (int64_t)conn[2 * e + k] + c];
            }
        int m = mat_id ? mat_id[e] : 0;
        double Emod = truss_modulus(kind[m], params[2 * m], params[2 * m + 1]);
        double se[2];
        for (int a = 0; a < 2; ++a) {
            double rec[32];
            if (dim == 3) truss_pair_d<3>(strain_model, X, U, Emod, area[e], a, rec, se);
            else if (dim == 2) truss_pair_d<2>(strain_model, X, U, Emod, area[e], a, rec, se);
            else truss_pair_d<1>(strain_model, X, U, Emod, area[e], a, rec, se);
            for (int r = 0; r < dim; ++r) {
                f[N * e + dim * a + r] = rec[2 * dim * dim + r];
                for (int b = 0; b < 2; ++b)
                    for (int q = 0; q < dim; ++q) K[N * N * e + (dim * a + r) + N * (dim * b + q)] = rec[b * dim * dim + dim * r + q];
            }
        }
        for (int k = 0; k < 9; ++k) sig[9 * e + k] = eps[9 * e + k] = 0.0;
        sig[9 * e] = se[0];
        eps[9 * e] = se[1];
    }
    return 0;
}

struct HsModel {
    MeshTables tab;
    std::vector<double> val, Fint, tet_out, truss_out;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> colidx;
    std::string err;
};

HsModel* hs_create(int dim, int64_t n_nodes, int64_t n_rows, int64_t n_tets, const int32_t* tets, int64_t n_trusses,
                   const int32_t* trusses) {
    HsModel* m = new HsModel();
    m->err = build_mesh_tables(dim, n_nodes, n_rows, n_tets, tets, n_trusses, trusses, m->tab);
    if (m->err.empty()) {
        m->val.assign((size_t)m->tab.n_slots() * dim * dim, 0.0);
        m->Fint.assign((size_t)n_rows * dim, 0.0);
        m->tet_out.assign((size_t)n_tets * 16, 0.0);
        m->truss_out.assign((size_t)n_trusses * 2, 0.0);
        bsell_to_csr_pattern(m->tab, m->rowptr, m->colidx);
    }
    return m;
}
const char* hs_error(HsModel* m) { return m->err.c_str(); }
void hs_destroy(HsModel* m) { delete m; }
// FNV-1a over every array and scalar of the host-built tables: a change of the builder (tables.cpp) that is meant to keep its
// output is checked against the recorded hashes (tests/golden/table_hashes.json)
uint64_t hs_tables_hash(HsModel* m) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) {
            h ^= b[i];
            h *= 1099511628211ull;
        }
    };
    auto vec = [&](const auto& v) {
        const uint64_t n = v.size();
        mix(&n, sizeof n);
        if (n) mix(v.data(), n * sizeof(v[0]));
    };
    const MeshTables& t = m->tab;
    const int64_t sc[6] = {t.dim, t.n_nodes, t.n_rows, t.n_slices, t.max_width, t.nnz_blocks};
    mix(sc, sizeof sc);
    vec(t.slice_ptr);
    vec(t.col);
    vec(t.row_nblk);
    for (int f = 0; f < 2; ++f) {
        const FamilyTables& F = t.fam[f];
        const int64_t fs[5] = {F.npe, F.rec, F.n_elem, F.max_snodes, F.max_pairs_per_slice};
        mix(fs, sizeof fs);
        vec(F.pair_ptr);
        vec(F.pair_code);
        vec(F.pair_lnodes);
        vec(F.snodes);
        vec(F.cptr);
        vec(F.ccode);
        const uint64_t nh = F.hdr.size();
        mix(&nh, sizeof nh);
        for (const SliceHdr& H : F.hdr) {  // field by field: the struct has no padding today, but a hash must not depend on it
            mix(&H.pair_base, sizeof H.pair_base);
            mix(&H.slot_base, sizeof H.slot_base);
            mix(&H.n_pairs, sizeof H.n_pairs);
            mix(&H.width, sizeof H.width);
            mix(H.row_off, sizeof H.row_off);
            mix(&H.n_snodes, sizeof H.n_snodes);
            mix(&H.snode_base, sizeof H.snode_base);
        }
    }
    return h;
}
int64_t hs_nnz(HsModel* m) { return (int64_t)m->colidx.size(); }
int64_t hs_stat(HsModel* m, int which) {
    switch (which) {
        case 0: return m->tab.n_slices;
        case 1: return m->tab.n_slots();
        case 2: return m->tab.nnz_blocks;
        case 3: return m->tab.fam[0].max_pairs_per_slice;
        case 4: return m->tab.fam[1].max_pairs_per_slice;
        default: return -1;
    }
}

// invariants of the slice node lists the assembly kernel stages in shared memory (tables.cpp): out = {violations, lists that are
// not strictly ascending, pairs whose local indices do not map back to their element's nodes, elements whose record is written
// by no pair or by more than one (among the elements that touch an owned node), largest list, total list entries}
void hs_check_slice_nodes(HsModel* m, int family, const int32_t* conn, int64_t n_elem, int64_t* out /*[6]*/) {
    const MeshTables& t = m->tab;
    const FamilyTables& F = t.fam[family];
    for (int k = 0; k < 6; ++k) out[k] = 0;
    if (F.n_elem == 0) return;
    std::vector<int> writers((size_t)n_elem, 0), touched((size_t)n_elem, 0);
    for (int64_t sl = 0; sl < t.n_slices; ++sl) {
        const SliceHdr& H = F.hdr[sl];
        const int32_t* sn = F.snodes.data() + H.snode_base;
        out[4] = std::max<int64_t>(out[4], H.n_snodes);
        out[5] += H.n_snodes;
        for (int k = 1; k < H.n_snodes; ++k)
            if (sn[k] <= sn[k - 1]) out[1]++;
        for (int tt = 0; tt < H.n_pairs; ++tt) {
            const int64_t p = H.pair_base + tt;
            const int64_t e = F.pair_code[p] / F.npe;
            touched[e] = 1;
            for (int b = 0; b < F.npe; ++b) {
                const uint16_t li = F.pair_lnodes[(size_t)p * F.npe + b] & 0x7fffu;
                if (li >= H.n_snodes || sn[li] != conn[e * F.npe + b]) out[2]++;
            }
            if (F.pair_lnodes[(size_t)p * F.npe] & 0x8000u) writers[e]++;
        }
    }
    for (int64_t e = 0; e < n_elem; ++e)
        if (touched[e] && writers[e] != 1) out[3]++;
    out[0] = out[1] + out[2] + out[3];
}

// slice ranges of the pipelined host-buffer assembly (tables.cpp::host_range_plan); returns the number of ranges
int hs_host_plan(HsModel* m, int chunks, int mid_weight, int64_t* slice0 /*[chunks+1]*/, int64_t* node_hi /*[chunks]*/) {
    std::vector<int64_t> s0, hi;
    host_range_plan(m->tab, chunks, mid_weight, s0, hi);
    for (size_t k = 0; k < s0.size(); ++k) slice0[k] = s0[k];
    for (size_t k = 0; k < hi.size(); ++k) node_hi[k] = hi[k];
    return (int)hi.size();
}

// walk k_assemble for both families at displacement Uv; `threads` emulates blockDim.x
int hs_assemble(HsModel* m, const int32_t* tets, const int32_t* tet_mat, const int32_t* trusses, const int32_t* truss_mat,
                const double* area, int strain_model, const int32_t* kind, const double* params, const double* xyz,
                const double* Uv, int threads) {
    const MeshTables& t = m->tab;
    const int dim = t.dim, BB = dim * dim;
    bool wrote = false;
    for (int fam = 0; fam < 2; ++fam) {
        const FamilyTables& F = t.fam[fam];
        if (F.n_elem == 0) continue;
        const int REC = fam == 0 ? TET_REC : truss_rec(dim);
        const int FOFF = fam == 0 ? 36 : 2 * BB;
        const bool ACCUM = wrote;
        std::vector<double> stage((size_t)std::max(F.max_pairs_per_slice, 1) * REC + ROW_SKEW * C);
        const int NPE = F.npe;
        std::vector<uint16_t> scode((size_t)std::max(F.max_pairs_per_slice, 1) * NPE), scp((size_t)t.max_width * C + 1);
        for (int64_t sl = 0; sl < t.n_slices; ++sl) {  // one CTA per slice
            const SliceHdr& H = F.hdr[sl];
            const int64_t p0 = H.pair_base, base = H.slot_base;
            const int np = H.n_pairs, width = H.width;
            const uint32_t cbase = (uint32_t)(p0 * NPE);
            const int nscp = width * C + 1;
            for (int tid = 0; tid < threads; ++tid)  // level-2 staging of the slot ranges
                for (int i = tid; i < nscp; i += threads) scp[i] = (uint16_t)(F.cptr[base * C + i] - cbase);
            for (int tid = 0; tid < threads; ++tid)  // phase A
                for (int tt = tid; tt < np; tt += threads) {
                    for (int b = 0; b < NPE; ++b) scode[(size_t)tt * NPE + b] = F.ccode[cbase + (size_t)tt * NPE + b];
                    const int32_t code = F.pair_code[p0 + tt];
                    int32_t nd[4];  // the element's nodes through the slice node list, as the kernel addresses them
                    for (int b = 0; b < NPE; ++b) nd[b] = F.snodes[(size_t)H.snode_base + (F.pair_lnodes[(size_t)(p0 + tt) * NPE + b] & 0x7fffu)];
                    const bool writer = (F.pair_lnodes[(size_t)(p0 + tt) * NPE] & 0x8000u) != 0;
                    int lrow = 0;
                    for (int k = 1; k < C; ++k) lrow += (tt >= H.row_off[k]) ? 1 : 0;
                    double* rec = stage.data() + (size_t)tt * REC + row_skew(fam, lrow);
                    if (fam == 0) {
                        const int64_t e = code >> 2;
                        const int a = code & 3;
                        double X[4][3], U[4][3];
                        for (int k = 0; k < 4; ++k)
                            for (int c = 0; c < 3; ++c) {
                                X[k][c] = xyz[3 * (int64_t)nd[k] + c];
                                U[k][c] = Uv[3 * (int64_t)nd[k] + c];
                            }
                        const int mm = tet_mat ? tet_mat[e] : 0;
                        tet_pair(kind[mm], X, U, params[2 * mm], params[2 * mm + 1], a, rec,
                                 writer ? m->tet_out.data() + 16 * e : nullptr);
                    } else {
                        const int64_t e = code >> 1;
                        const int a = code & 1;
                        double X[2][3] = {{0}}, U[2][3] = {{0}};
                        for (int k = 0; k < 2; ++k)
                            for (int c = 0; c < dim; ++c) {
                                X[k][c] = xyz[dim * (int64_t)nd[k] + c];
                                U[k][c] = Uv[dim * (int64_t)nd[k] + c];
                            }
                        const int mm = truss_mat ? truss_mat[e] : 0;
                        const double Emod = truss_modulus(kind[mm], params[2 * mm], params[2 * mm + 1]);
                        double se[2];
                        if (dim == 3) truss_pair_d<3>(strain_model, X, U, Emod, area[e], a, rec, se);
                        else if (dim == 2) truss_pair_d<2>(strain_model, X, U, Emod, area[e], a, rec, se);
                        else truss_pair_d<1>(strain_model, X, U, Emod, area[e], a, rec, se);
                        if (writer) {
                            m->truss_out[2 * e] = se[0];
                            m->truss_out[2 * e + 1] = se[1];
                        }
                    }
                }
            // phase B: one item per (slot, block row r)
            const int nK = width * dim * C, nF = C * dim;
            double* vout = m->val.data() + base * BB * C;
            for (int tid = 0; tid < threads; ++tid)
                for (int w = tid; w < nK + nF; w += threads) {
                    if (w < nK) {
                        const int lane = w % C, r = (w / C) % dim, s = w / (C * dim);
                        const int slot = s * C + lane;
                        double acc[3] = {0, 0, 0};
                        for (int q = scp[slot]; q < scp[slot + 1]; ++q)
                            for (int j = 0; j < dim; ++j) acc[j] += stage[(size_t)scode[q] + r * dim + j];
                        double* dst = vout + ((size_t)s * BB + r * dim) * C + lane;
                        for (int j = 0; j < dim; ++j) {
                            if (ACCUM) acc[j] += dst[j * C];
                            dst[j * C] = acc[j];
                        }
                    } else {
                        const int j = w - nK, lane = j / dim, r = j % dim;
                        const int t0 = H.row_off[lane], t1 = H.row_off[lane + 1];
                        const int64_t row = sl * C + lane;
                        if (t1 > t0 || !ACCUM) {
                            double acc = 0.0;
                            for (int tt = t0; tt < t1; ++tt) acc += stage[(size_t)tt * REC + row_skew(fam, lane) + FOFF + r];
                            if (row < t.n_rows) {
                                if (ACCUM) acc += m->Fint[row * dim + r];
                                m->Fint[row * dim + r] = acc;
                            }
                        }
                    }
                }
        }
        wrote = true;
    }
    return 0;
}

void hs_get(HsModel* m, int64_t* rowptr, int32_t* col, double* csr_val, double* Fint, double* tet_out, double* truss_out) {
    std::copy(m->rowptr.begin(), m->rowptr.end(), rowptr);
    std::copy(m->colidx.begin(), m->colidx.end(), col);
    bsell_to_csr_values(m->tab, m->val.data(), csr_val);
    std::copy(m->Fint.begin(), m->Fint.end(), Fint);
    std::copy(m->tet_out.begin(), m->tet_out.end(), tet_out);
    std::copy(m->truss_out.begin(), m->truss_out.end(), truss_out);
}

// spmv_row walk: y = M K M x on the BSELL arrays (x has n_nodes*dim entries)
void hs_spmv(HsModel* m, const uint8_t* mask, const double* x, double* y) {
    const MeshTables& t = m->tab;
    const int BS = t.dim;
    for (int64_t row = 0; row < t.n_rows; ++row) {
        const int64_t sl = row / C;
        const int lane = (int)(row % C);
        const int64_t base = t.slice_ptr[sl];
        const int width = (int)(t.slice_ptr[sl + 1] - base);
        double acc[3] = {0, 0, 0};
        for (int s = 0; s < width; ++s) {
            const int64_t cn = t.col[(base + s) * C + lane];
            for (int r = 0; r < BS; ++r)
                for (int q = 0; q < BS; ++q)
                    acc[r] += m->val[((base + s) * BS * BS + r * BS + q) * C + lane] * x[cn * BS + q];
        }
        for (int r = 0; r < BS; ++r) y[row * BS + r] = mask[row * BS + r] ? acc[r] : 0.0;
    }
}

// PCG in the device phase order (prologue, [update_p, spmv+dot, update_xr]*, epilogue)
int hs_pcg(HsModel* m, const uint8_t* mask, const double* b, double* x, int precond, double reltol, double abstol,
           int64_t maxiter, int64_t* iters, double* res_out) {
    const MeshTables& t = m->tab;
    const int BS = t.dim;
    const int64_t n = t.n_rows * BS;
    std::vector<double> r(n), p((size_t)t.n_nodes * BS, 0.0), Ap(n), dinv(n);
    double rr = 0, rho = 0;
    for (int64_t i = 0; i < n; ++i) {
        const bool fr = mask[i] != 0;
        const double ri = fr ? b[i] : 0.0;
        double di = fr ? 1.0 : 0.0;
        if (fr && precond) {
            const int64_t row = i / BS;
            const int c = (int)(i % BS);
            const int64_t base = t.slice_ptr[row / C];
            int s = 0;
            while (t.col[(base + s) * C + row % C] != row) ++s;
            di = 1.0 / m->val[((base + s) * BS * BS + c * BS + c) * C + (row % C)];
        }
        r[i] = ri;
        x[i] = 0.0;
        dinv[i] = di;
        rr += ri * ri;
        rho += ri * ri * di;
    }
    double res = std::sqrt(rr);
    const double tol = std::max(reltol * res, abstol);
    double rho_prev = 1.0;
    int64_t it = 0;
    while (!(it >= maxiter || res <= tol)) {
        const double beta = rho / rho_prev;
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] * dinv[i] + beta * p[i];
        hs_spmv(m, mask, p.data(), Ap.data());
        double pAp = 0;
        for (int64_t i = 0; i < n; ++i) pAp += p[i] * Ap[i];
        const double alpha = rho / pAp;
        rr = 0;
        double rz = 0;
        for (int64_t i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * Ap[i];
            rr += r[i] * r[i];
            rz += r[i] * r[i] * dinv[i];
        }
        rho_prev = rho;
        rho = rz;
        res = std::sqrt(rr);
        ++it;
    }
    *iters = it;
    *res_out = res;
    return 0;
}

}  // extern "C"
