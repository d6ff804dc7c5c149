// tables.cpp -- see tables.hpp.  Plain C++ (no CUDA) so tests can exercise it without a GPU.
#include "tables.hpp"

#include <algorithm>
#include <cstring>

namespace onsas {

namespace {

// sorted unique neighbour nodes of row i gathered from the pairs of both families
inline int gather_row_cols(const MeshTables& t, const int32_t* const conn[2], int64_t i, std::vector<int32_t>& buf) {
    buf.clear();
    for (int f = 0; f < 2; ++f) {
        const FamilyTables& F = t.fam[f];
        if (F.n_elem == 0) continue;
        for (int64_t p = F.pair_ptr[i]; p < F.pair_ptr[i + 1]; ++p) {
            int64_t e = F.pair_code[p] / F.npe;
            for (int b = 0; b < F.npe; ++b) buf.push_back(conn[f][e * F.npe + b]);
        }
    }
    buf.push_back((int32_t)i);  // a row always owns its diagonal block (isolated nodes included)
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
    return (int)buf.size();
}

}  // namespace

std::string build_mesh_tables(int dim, int64_t n_nodes, int64_t n_rows, int64_t n_tets, const int32_t* tets,
                              int64_t n_trusses, const int32_t* trusses, MeshTables& t) {
    constexpr int C = SLICE_ROWS;
    if (dim < 1 || dim > 3) return "dim must be 1, 2 or 3";
    if (n_tets > 0 && dim != 3) return "tetrahedra need dim == 3";
    if (n_rows < 0 || n_rows > n_nodes) return "owned node count out of range";
    t = MeshTables();
    t.dim = dim;
    t.n_nodes = n_nodes;
    t.n_rows = n_rows;
    const int32_t* conn[2] = {tets, trusses};
    const int64_t ne[2] = {n_tets, n_trusses};
    const int npe[2] = {4, 2};

    // ---- 1. pair lists (counting sort by row keeps element order)
    for (int f = 0; f < 2; ++f) {
        FamilyTables& F = t.fam[f];
        F.npe = npe[f];
        F.rec = f == 0 ? TET_REC : truss_rec(dim);
        F.n_elem = ne[f];
        F.pair_ptr.assign(n_rows + 1, 0);
        if (ne[f] == 0) continue;
        if (ne[f] * npe[f] > (int64_t)0x7fffffff) return "too many elements for 32-bit pair codes";
        for (int64_t q = 0; q < ne[f] * npe[f]; ++q) {
            int32_t nd = conn[f][q];
            if (nd < 0 || nd >= n_nodes) return "element references a node id out of range";
            if (nd < n_rows) F.pair_ptr[nd + 1]++;
        }
        for (int64_t i = 0; i < n_rows; ++i) F.pair_ptr[i + 1] += F.pair_ptr[i];
        F.pair_code.resize(F.pair_ptr[n_rows]);
        std::vector<int64_t> fill(F.pair_ptr.begin(), F.pair_ptr.end() - 1);
        for (int64_t q = 0; q < ne[f] * npe[f]; ++q) {
            int32_t nd = conn[f][q];
            if (nd < n_rows) F.pair_code[fill[nd]++] = (int32_t)q;  // q = e*npe + a
        }
    }

    // ---- 2. block pattern: count, slice widths, fill
    t.row_nblk.assign(n_rows, 0);
#pragma omp parallel
    {
        std::vector<int32_t> buf;
        buf.reserve(256);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n_rows; ++i) t.row_nblk[i] = gather_row_cols(t, conn, i, buf);
    }
    t.n_slices = (n_rows + C - 1) / C;
    t.slice_ptr.assign(t.n_slices + 1, 0);
    t.nnz_blocks = 0;
    for (int64_t sl = 0; sl < t.n_slices; ++sl) {
        int w = 0;
        for (int l = 0; l < C; ++l) {
            int64_t i = sl * C + l;
            if (i < n_rows) {
                w = std::max(w, t.row_nblk[i]);
                t.nnz_blocks += t.row_nblk[i];
            }
        }
        t.slice_ptr[sl + 1] = t.slice_ptr[sl] + w;
        t.max_width = std::max(t.max_width, w);
    }
    if (t.slice_ptr[t.n_slices] * C * (int64_t)dim * dim > (int64_t)0x7fffffff * 8) return "matrix too large";
    t.col.assign((size_t)t.slice_ptr[t.n_slices] * C, 0);
#pragma omp parallel
    {
        std::vector<int32_t> buf;
        buf.reserve(256);
#pragma omp for schedule(static)
        for (int64_t sl = 0; sl < t.n_slices; ++sl) {
            int64_t base = t.slice_ptr[sl];
            int w = (int)(t.slice_ptr[sl + 1] - base);
            for (int l = 0; l < C; ++l) {
                int64_t i = sl * C + l;
                int nb = 0;
                if (i < n_rows) nb = gather_row_cols(t, conn, i, buf);
                int32_t pad = (int32_t)(i < n_rows ? i : 0);
                for (int s = 0; s < w; ++s) t.col[(base + s) * C + l] = s < nb ? buf[s] : pad;
            }
        }
    }

    // ---- 3. contribution lists per family
    for (int f = 0; f < 2; ++f) {
        FamilyTables& F = t.fam[f];
        if (F.n_elem == 0) continue;
        int64_t n_slots = t.n_slots();
        int64_t n_contrib = (int64_t)F.pair_code.size() * F.npe;
        if (n_contrib >= (int64_t)0xffffffffu) return "too many contributions for 32-bit offsets";
        F.cptr.assign(n_slots + 1, 0);
        F.ccode.assign(n_contrib, 0);
        F.hdr.assign(t.n_slices, SliceHdr());
        // slice node lists: the distinct nodes the elements of a slice touch.  The assembly CTA stages X and U of exactly
        // these nodes in shared memory once (runs of consecutive ids -> coalesced loads) and its pairs address them with
        // 16-bit local indices, instead of every pair gathering its element's nodes from global memory.
        std::vector<uint32_t> sn_ptr((size_t)t.n_slices + 1, 0);
        {
            int32_t max_sn = 0;
#pragma omp parallel
            {
                std::vector<int32_t> buf;
#pragma omp for schedule(static) reduction(max : max_sn)
                for (int64_t sl = 0; sl < t.n_slices; ++sl) {
                    const int64_t r0 = sl * C, r1 = std::min<int64_t>(r0 + C, n_rows);
                    buf.clear();
                    for (int64_t p = F.pair_ptr[r0]; p < F.pair_ptr[r1]; ++p) {
                        const int64_t e = F.pair_code[p] / F.npe;
                        for (int b = 0; b < F.npe; ++b) buf.push_back(conn[f][e * F.npe + b]);
                    }
                    std::sort(buf.begin(), buf.end());
                    const int32_t n = (int32_t)(std::unique(buf.begin(), buf.end()) - buf.begin());
                    sn_ptr[sl + 1] = (uint32_t)n;
                    max_sn = std::max(max_sn, n);
                }
            }
            if (max_sn > 0x7fff) return "too many distinct nodes in one 8-row slice (node valence too high)";
            uint64_t tot = 0;
            for (int64_t sl = 0; sl < t.n_slices; ++sl) {
                tot += sn_ptr[sl + 1];
                sn_ptr[sl + 1] = (uint32_t)tot;
            }
            if (tot >= 0xffffffffull) return "slice node lists exceed 32-bit offsets";
            F.max_snodes = max_sn;
            F.snodes.assign((size_t)tot, 0);
            F.pair_lnodes.assign((size_t)n_contrib, 0);
        }
        const int BB = dim * dim;
        int32_t max_pairs = 0;
        bool overflow = false;
#pragma omp parallel for schedule(static) reduction(max : max_pairs) reduction(|| : overflow)
        for (int64_t sl = 0; sl < t.n_slices; ++sl) {
            int64_t r0 = sl * C, r1 = std::min<int64_t>(r0 + C, n_rows);
            int64_t p0 = F.pair_ptr[r0], p1 = F.pair_ptr[r1];
            int64_t base = t.slice_ptr[sl];
            int w = (int)(t.slice_ptr[sl + 1] - base);
            int32_t np = (int32_t)(p1 - p0);
            max_pairs = std::max(max_pairs, np);
            if ((int64_t)np * F.rec + ROW_SKEW * C > 65535) {  // contribution codes are 16-bit shared-memory offsets
                overflow = true;
                continue;
            }
            SliceHdr& H = F.hdr[sl];
            H.pair_base = p0;
            H.slot_base = base;
            H.n_pairs = np;
            H.width = w;
            {
                int32_t* sn = F.snodes.data() + sn_ptr[sl];
                const int32_t nsn = (int32_t)(sn_ptr[sl + 1] - sn_ptr[sl]);
                H.snode_base = sn_ptr[sl];
                H.n_snodes = (uint16_t)nsn;
                (void)nsn;
                std::vector<int32_t> buf;
                buf.reserve((size_t)np * F.npe);
                for (int64_t p = p0; p < p1; ++p) {
                    const int64_t e = F.pair_code[p] / F.npe;
                    for (int b = 0; b < F.npe; ++b) buf.push_back(conn[f][e * F.npe + b]);
                }
                std::sort(buf.begin(), buf.end());
                buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
                std::copy(buf.begin(), buf.end(), sn);
                for (int64_t p = p0; p < p1; ++p) {
                    const int64_t e = F.pair_code[p] / F.npe;
                    const int a = (int)(F.pair_code[p] % F.npe);
                    int first_owned = 0;
                    while (first_owned < F.npe && conn[f][e * F.npe + first_owned] >= n_rows) ++first_owned;
                    for (int b = 0; b < F.npe; ++b) {
                        const int32_t nd = conn[f][e * F.npe + b];
                        uint16_t li = (uint16_t)(std::lower_bound(buf.begin(), buf.end(), nd) - buf.begin());
                        if (b == 0 && a == first_owned) li |= 0x8000u;
                        F.pair_lnodes[(size_t)p * F.npe + b] = li;
                    }
                }
            }
            for (int l = 0; l <= C; ++l) H.row_off[l] = (uint16_t)(F.pair_ptr[std::min<int64_t>(r0 + l, r1)] - p0);
            std::vector<uint32_t> cnt((size_t)w * C + 1, 0);
            std::vector<uint16_t> slot_of((size_t)np * F.npe);
            for (int64_t i = r0; i < r1; ++i) {
                int l = (int)(i - r0);
                int nb = t.row_nblk[i];
                for (int64_t p = F.pair_ptr[i]; p < F.pair_ptr[i + 1]; ++p) {
                    int64_t e = F.pair_code[p] / F.npe;
                    for (int b = 0; b < F.npe; ++b) {
                        int32_t target = conn[f][e * F.npe + b];
                        int lo = 0, hi = nb - 1, s = -1;  // binary search in the row's sorted block list
                        while (lo <= hi) {
                            int mid = (lo + hi) >> 1;
                            int32_t c = t.col[(base + mid) * C + l];
                            if (c < target)
                                lo = mid + 1;
                            else if (c > target)
                                hi = mid - 1;
                            else {
                                s = mid;
                                break;
                            }
                        }
                        uint16_t slot = (uint16_t)(s * C + l);
                        slot_of[(p - p0) * F.npe + b] = slot;
                        cnt[slot + 1]++;
                    }
                }
            }
            for (size_t k = 0; k < (size_t)w * C; ++k) cnt[k + 1] += cnt[k];
            uint32_t gbase = (uint32_t)(p0 * F.npe);
            for (size_t k = 0; k <= (size_t)w * C; ++k) F.cptr[base * C + k] = gbase + cnt[k];
            std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
            // record of local pair lp (row l) sits at lp*rec + row_skew(family, l): the skew keeps the rows of a structured mesh
            // (24 pairs * 39 doubles = 8 mod 16 bank pairs apart) on different shared-memory banks (tables.hpp)
            for (int64_t i = r0; i < r1; ++i)
                for (int64_t p = F.pair_ptr[i]; p < F.pair_ptr[i + 1]; ++p) {
                    const int32_t lp = (int32_t)(p - p0);
                    for (int b = 0; b < F.npe; ++b) {
                        uint16_t slot = slot_of[(size_t)lp * F.npe + b];
                        F.ccode[gbase + fill[slot]++] = (uint16_t)(lp * F.rec + row_skew(f, (int)(i - r0)) + b * BB);
                    }
                }
        }
        if (overflow) return "too many element pairs in one 8-row slice (node valence too high)";
        F.max_pairs_per_slice = max_pairs;
        F.cptr[n_slots] = (uint32_t)n_contrib;
    }
    return std::string();
}

void bsell_to_csr_pattern(const MeshTables& t, std::vector<int64_t>& rowptr, std::vector<int32_t>& colidx) {
    constexpr int C = SLICE_ROWS;
    const int d = t.dim;
    rowptr.assign(t.n_rows * d + 1, 0);
    for (int64_t i = 0; i < t.n_rows; ++i)
        for (int r = 0; r < d; ++r) rowptr[i * d + r + 1] = (int64_t)t.row_nblk[i] * d;
    for (int64_t k = 0; k < t.n_rows * d; ++k) rowptr[k + 1] += rowptr[k];
    colidx.resize(rowptr[t.n_rows * d]);
    for (int64_t i = 0; i < t.n_rows; ++i) {
        int64_t sl = i / C;
        int l = (int)(i % C);
        int64_t base = t.slice_ptr[sl];
        for (int r = 0; r < d; ++r) {
            int64_t o = rowptr[i * d + r];
            for (int s = 0; s < t.row_nblk[i]; ++s) {
                int32_t c = t.col[(base + s) * C + l];
                for (int q = 0; q < d; ++q) colidx[o++] = c * d + q;
            }
        }
    }
}

void bsell_to_csr_values(const MeshTables& t, const double* val, double* csr_val) {
    constexpr int C = SLICE_ROWS;
    const int d = t.dim;
    int64_t o = 0;
    for (int64_t i = 0; i < t.n_rows; ++i) {
        int64_t sl = i / C;
        int l = (int)(i % C);
        int64_t base = t.slice_ptr[sl];
        for (int r = 0; r < d; ++r)
            for (int s = 0; s < t.row_nblk[i]; ++s)
                for (int q = 0; q < d; ++q) csr_val[o++] = val[((base + s) * d * d + (r * d + q)) * C + l];
    }
}

void host_range_plan(const MeshTables& t, int chunks, int mid_weight, std::vector<int64_t>& slice0, std::vector<int64_t>& node_hi) {
    const int64_t ns = t.n_slices;
    const int nch = (int)std::max<int64_t>(1, std::min<int64_t>(std::max(1, chunks), ns));
    // the first and the last range are short: the first kernel waits for its piece of U and the last piece of F_int
    // leaves after the last kernel -- the two exposed copies
    slice0.assign((size_t)nch + 1, 0);
    const int wm = std::max(1, mid_weight);
    const int64_t wt = nch <= 2 ? nch : 2 + (int64_t)(nch - 2) * wm;
    int64_t acc = 0;
    for (int k = 0; k < nch; ++k) {
        acc += (nch <= 2 || k == 0 || k == nch - 1) ? 1 : wm;
        slice0[k + 1] = ns * acc / wt;
    }
    node_hi.assign((size_t)nch, 0);
    for (int k = 0; k < nch; ++k) {
        int64_t hi = std::min<int64_t>(slice0[k + 1] * SLICE_ROWS, t.n_rows);  // the rows themselves
        for (int f = 0; f < 2; ++f) {
            const FamilyTables& T = t.fam[f];
            if (T.n_elem == 0 || T.hdr.empty() || slice0[k + 1] == slice0[k]) continue;
            int32_t mx = -1;  // the slice node lists are ascending: the last entry of each is its largest node id
            for (int64_t sl = slice0[k]; sl < slice0[k + 1]; ++sl) {
                const SliceHdr& h = T.hdr[(size_t)sl];
                // halo nodes (ids >= n_rows, multi-GPU) are not part of the prefix: their values travel first, as one block
                const int32_t* b = T.snodes.data() + h.snode_base;
                const int32_t* e = std::lower_bound(b, b + h.n_snodes, (int32_t)std::min<int64_t>(t.n_rows, 0x7fffffff));
                if (e > b) mx = std::max(mx, e[-1]);
            }
            hi = std::max<int64_t>(hi, (int64_t)mx + 1);
        }
        node_hi[k] = std::max(hi, k > 0 ? node_hi[k - 1] : 0);
    }
}

}  // namespace onsas
