// Bank model of the assembly kernel's shared-memory traffic (DESIGN.md section 3.1, tables.hpp row_skew): wavefronts of
// the phase A stores (one STS.64 per record entry and pair) and of the phase B loads (three LDS.64 per contribution of
// every (slot, row) item) on interior slices of the structured meshes, for a record stride REC and per-row skews.
//   model: a warp-wide 64-bit access costs max over the 32 four-byte banks of the distinct words it touches there.
// usage: bank_model tet|truss [cells] [REC] [skew of row 0 .. 7]      (CPU only; links the product's table builder)
//   ./bank_model tet 23 39 0 0 2 2 4 4 6 6     -> phase A 468, phase B 537   (shipped)
//   ./bank_model tet 23 39 0 1 2 3 4 5 6 7     -> 624 / 669                  (one double per row, r03)
//   ./bank_model truss 23 21 0 0 0 0 0 0 0 0   -> 147 / 147                  (shipped)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include "tables.hpp"
using namespace onsas;

static int wavefronts(const std::vector<long>& words) {
    std::set<long> per_bank[32];
    for (long a : words) per_bank[a & 31].insert(a);
    size_t m = 0;
    for (auto& s : per_bank) m = std::max(m, s.size());
    return (int)m;
}

int main(int argc, char** argv) {
    const bool truss = argc > 1 && !strcmp(argv[1], "truss");
    const int fam = truss ? 1 : 0;
    const int n = argc > 2 ? atoi(argv[2]) : 23;
    const int REC = argc > 3 ? atoi(argv[3]) : (truss ? truss_rec(3) : TET_REC);
    int off[8];
    for (int l = 0; l < 8; ++l) off[l] = argc > 4 + l ? atoi(argv[4 + l]) : row_skew(fam, l);
    const int nn = n + 1;
    auto id = [&](int i, int j, int k) { return i + nn * (j + nn * k); };
    std::vector<int32_t> conn;
    if (!truss) {  // the reference's 6-tet split of every hexahedron (uniaxial_extension.jl:45-72)
        const int cor[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 1}, {0, 1, 0}, {1, 0, 0}, {1, 0, 1}, {1, 1, 1}, {1, 1, 0}};
        const int tt[6][4] = {{1, 4, 2, 6}, {6, 2, 3, 4}, {4, 3, 6, 7}, {4, 1, 5, 6}, {4, 6, 5, 8}, {4, 7, 6, 8}};
        for (int k = 0; k < n; ++k)
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i)
                    for (auto& t : tt)
                        for (int a = 0; a < 4; ++a) conn.push_back(id(i + cor[t[a] - 1][0], j + cor[t[a] - 1][1], k + cor[t[a] - 1][2]));
    } else {  // braced lattice of meshgen.truss_lattice: axis bars, face diagonals, body diagonal, sorted by first node
        const int dirs[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
        std::vector<std::pair<int, int>> bl;
        for (auto& d : dirs)
            for (int k = 0; k + d[2] < nn; ++k)
                for (int j = 0; j + d[1] < nn; ++j)
                    for (int i = 0; i + d[0] < nn; ++i) bl.push_back({id(i, j, k), id(i + d[0], j + d[1], k + d[2])});
        std::sort(bl.begin(), bl.end());
        for (auto& b : bl) {
            conn.push_back(b.first);
            conn.push_back(b.second);
        }
    }
    MeshTables T;
    const int64_t nnode = (int64_t)nn * nn * nn;
    std::string e = truss ? build_mesh_tables(3, nnode, nnode, 0, nullptr, (int64_t)conn.size() / 2, conn.data(), T)
                          : build_mesh_tables(3, nnode, nnode, (int64_t)conn.size() / 4, conn.data(), 0, nullptr, T);
    if (!e.empty()) {
        printf("error: %s\n", e.c_str());
        return 1;
    }
    const FamilyTables& F = T.fam[fam];
    const int NPE = F.npe, rec_tab = F.rec, nent = rec_tab, full = truss ? 112 : 192, nth = truss ? 128 : 192;
    long totA = 0, totB = 0, idealB = 0, instrB = 0, nsl = 0;
    for (int64_t sl = 0; sl < T.n_slices && nsl < 8; ++sl) {
        const SliceHdr& h = F.hdr[sl];
        if (h.n_pairs != full) continue;  // interior slices only
        ++nsl;
        const int np = h.n_pairs;
        std::vector<int> rowof(np);
        for (int t = 0; t < np; ++t) {
            int l = 0;
            for (int k = 1; k < 8; ++k) l += t >= h.row_off[k];
            rowof[t] = l;
        }
        for (int w0 = 0; w0 < np; w0 += 32)  // phase A
            for (int k = 0; k < nent; ++k) {
                std::vector<long> words;
                for (int t = w0; t < std::min(np, w0 + 32); ++t) {
                    const long d = (long)t * REC + off[rowof[t]] + k;
                    words.push_back(2 * d);
                    words.push_back(2 * d + 1);
                }
                totA += wavefronts(words);
            }
        const uint32_t cbase = (uint32_t)(h.pair_base * NPE);
        const int nK = h.width * 3 * 8;
        for (int w0 = 0; w0 < nK; w0 += 32) {  // phase B: item w = lane + 8 (r + 3 s)
            std::vector<std::vector<long>> src(32);
            int maxc = 0;
            for (int w = w0; w < std::min(nK, w0 + 32); ++w) {
                const int lane = w % 8, r = (w / 8) % 3, s = w / 24;
                const int64_t slot = (h.slot_base + s) * 8 + lane;
                const uint32_t q0 = F.cptr[slot] - cbase, q1 = F.cptr[slot + 1] - cbase;
                for (uint32_t q = q0; q < q1; ++q) {
                    const int code = F.ccode[cbase + q];  // lp * rec + row_skew(family, l) + 9 b
                    const int lp = code / rec_tab, l = rowof[lp], b9 = code - lp * rec_tab - row_skew(fam, l);
                    src[w - w0].push_back((long)lp * REC + off[l] + b9 + r * 3);
                }
                maxc = std::max(maxc, (int)src[w - w0].size());
            }
            for (int i = 0; i < maxc; ++i)
                for (int j = 0; j < 3; ++j) {
                    std::vector<long> words;
                    int act = 0;
                    for (int ln = 0; ln < 32; ++ln)
                        if ((int)src[ln].size() > i) {
                            const long d = src[ln][i] + j;
                            words.push_back(2 * d);
                            words.push_back(2 * d + 1);
                            ++act;
                        }
                    totB += wavefronts(words);
                    idealB += (act + 15) / 16;
                    ++instrB;
                }
        }
        (void)nth;
    }
    if (nsl == 0) {
        printf("no interior slice with %d pairs (use more cells)\n", full);
        return 1;
    }
    printf("%s cells=%d REC=%d skew=%d,%d,%d,%d,%d,%d,%d,%d | per interior slice: phase A store wavefronts %.1f, phase B load wavefronts %.1f "
           "(lower bound %.1f, %.1f LDS.64 per warp-slot), total %.1f\n",
           truss ? "truss" : "tet", n, REC, off[0], off[1], off[2], off[3], off[4], off[5], off[6], off[7], (double)totA / nsl,
           (double)totB / nsl, (double)idealB / nsl, (double)instrB / nsl, (double)(totA + totB) / nsl);
    return 0;
}
