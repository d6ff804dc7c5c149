// partition.cpp -- see partition.hpp.  Plain C++ (no CUDA) so tests can exercise it without a GPU.
#include "partition.hpp"

#include <algorithm>
#include <cstring>

namespace onsas {

namespace {

// recursive coordinate bisection of ids[lo, hi) into `parts` parts; every finished part is left in ascending id order
void rcb(const double* xyz, int dim, std::vector<int32_t>& ids, size_t lo, size_t hi, int parts, std::vector<int64_t>& sizes) {
    if (parts <= 1) {
        std::sort(ids.begin() + lo, ids.begin() + hi);
        sizes.push_back((int64_t)(hi - lo));
        return;
    }
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) {
        double a = 1e300, b = -1e300;
#pragma omp parallel for reduction(min : a) reduction(max : b) if (hi - lo > 100000)
        for (int64_t k = (int64_t)lo; k < (int64_t)hi; ++k) {
            const double v = xyz[(size_t)ids[k] * dim + d];
            a = std::min(a, v);
            b = std::max(b, v);
        }
        mn[d] = a;
        mx[d] = b;
    }
    int ax = 0;  // first axis of largest extent
    for (int d = 1; d < dim; ++d)
        if (hi > lo && mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
    const int pl = parts / 2;
    const size_t mid = lo + ((hi - lo) * (size_t)pl) / (size_t)parts;
    std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int32_t a, int32_t b) {
        const double va = xyz[(size_t)a * dim + ax], vb = xyz[(size_t)b * dim + ax];
        return va < vb || (va == vb && a < b);  // ties broken by node id: deterministic
    });
    rcb(xyz, dim, ids, lo, mid, pl, sizes);
    rcb(xyz, dim, ids, mid, hi, parts - pl, sizes);
}

}  // namespace

int coarse_aggregate_count(int64_t n_owned_nodes, int dim) {
    const int cd = dim == 3 ? 6 : dim;  // coarse dofs per aggregate: translations (+ rotations in 3D)
    return (int)std::max<int64_t>(1, std::min<int64_t>(n_owned_nodes / COARSE_TARGET_NODES, COARSE_NC_MAX / cd));
}

int Partition::owner_of(int64_t new_id) const {
    return (int)(std::upper_bound(ranges.begin(), ranges.end(), new_id) - ranges.begin()) - 1;
}

namespace {
// Morton (Z-curve) key of a point inside a bounding box: 21 bits per axis, interleaved
inline uint64_t spread21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
// nodes ids[lo, hi) re-ordered along the Z-curve of their coordinates (ties by id): consecutive ids = spatial neighbours,
// whatever numbering the caller's mesh generator produced
void morton_sort(const double* xyz, int dim, std::vector<int32_t>& ids, size_t lo, size_t hi) {
    if (hi - lo < 2) return;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (size_t k = lo; k < hi; ++k)
        for (int d = 0; d < dim; ++d) {
            const double v = xyz[(size_t)ids[k] * dim + d];
            mn[d] = std::min(mn[d], v);
            mx[d] = std::max(mx[d], v);
        }
    double ext = 0.0;  // one scale for all axes: the curve follows cubes, not the box's aspect ratio
    for (int d = 0; d < dim; ++d) ext = std::max(ext, mx[d] - mn[d]);
    const double sc = ext > 0.0 ? 2097151.0 / ext : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> key(hi - lo);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)(hi - lo); ++k) {
        const int32_t id = ids[lo + (size_t)k];
        uint64_t m = 0;
        for (int d = 0; d < dim; ++d) m |= spread21((uint64_t)((xyz[(size_t)id * dim + d] - mn[d]) * sc)) << d;
        key[(size_t)k] = std::make_pair(m, id);
    }
    std::sort(key.begin(), key.end());
    for (size_t k = 0; k < key.size(); ++k) ids[lo + k] = key[k].second;
}
}  // namespace

std::string build_partition(int dim, int64_t n_nodes, const double* xyz, int64_t n_tets, const int32_t* tets,
                            const int32_t* tet_mat, int64_t n_trusses, const int32_t* trusses, const int32_t* truss_mat,
                            const double* area, int64_t n_free, const int64_t* free_dofs, int n_ranks, int reorder, Partition& P) {
    if (n_ranks < 1 || n_ranks > PART_MAX_RANKS) return "the partitioner supports 1 to 16 ranks";
    if (dim < 1 || dim > 3) return "dim must be 1, 2 or 3";
    if (n_nodes >= (int64_t)0x7fffffff) return "too many nodes for 32-bit node ids";
    P = Partition();
    P.n_ranks = n_ranks;
    P.dim = dim;
    P.n_nodes = n_nodes;
    P.n_tets = n_tets;
    P.n_trusses = n_trusses;
    // ---- 1. RCB of the nodes, renumbering
    P.order.resize((size_t)n_nodes);
    for (int64_t i = 0; i < n_nodes; ++i) P.order[i] = (int32_t)i;
    std::vector<int64_t> sizes;
    rcb(xyz, dim, P.order, 0, (size_t)n_nodes, n_ranks, sizes);
    P.ranges.assign((size_t)n_ranks + 1, 0);
    for (int p = 0; p < n_ranks; ++p) P.ranges[p + 1] = P.ranges[p] + sizes[p];
    if (reorder == 1)  // inside every part: Z-curve order instead of the caller's order
        for (int p = 0; p < n_ranks; ++p) morton_sort(xyz, dim, P.order, (size_t)P.ranges[p], (size_t)P.ranges[p + 1]);
    P.agg_ptr.assign((size_t)n_ranks, std::vector<int32_t>());
    if (reorder == 2) {
        // aggregate-major: every part is cut further (the same recursive coordinate bisection) into the node aggregates
        // of the two-level preconditioner, the nodes of an aggregate are numbered consecutively (Z-curve inside it).
        // The solver's aggregate-ordered passes (w = Z^T r with the residual update, z = D^-1 r + Z y) then stream
        // through memory instead of gathering, and an 8-row slice of K still holds spatial neighbours.
        for (int p = 0; p < n_ranks; ++p) {
            const int64_t lo = P.ranges[p], hi = P.ranges[p + 1];
            const int n_agg = coarse_aggregate_count(hi - lo, dim);
            std::vector<int64_t> asz;
            rcb(xyz, dim, P.order, (size_t)lo, (size_t)hi, n_agg, asz);
            std::vector<int32_t>& ap = P.agg_ptr[(size_t)p];
            ap.assign(1, 0);
            for (int64_t z : asz) ap.push_back(ap.back() + (int32_t)z);
            for (size_t a = 0; a + 1 < ap.size(); ++a) morton_sort(xyz, dim, P.order, (size_t)(lo + ap[a]), (size_t)(lo + ap[a + 1]));
        }
    }
    P.inv.resize((size_t)n_nodes);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_nodes; ++k) P.inv[P.order[k]] = (int32_t)k;
    if (n_ranks > 1 && reorder != 2) {
        // level-2 aggregates: bisect a COPY of every part's node list (the numbering stays as it is)
        const int cd = dim == 3 ? 6 : dim;
        int64_t smallest = n_nodes;
        for (int p = 0; p < n_ranks; ++p) smallest = std::min(smallest, P.ranges[p + 1] - P.ranges[p]);
        const int n2 = (int)std::max<int64_t>(1, std::min<int64_t>((COARSE_NC_MAX / cd) / n_ranks, smallest / 64));
        P.n_agg2_per_rank = n2;
        P.agg2.assign((size_t)n_nodes, 0);
        P.cen2.assign((size_t)n2 * n_ranks * dim, 0.0);
        for (int p = 0; p < n_ranks; ++p) {
            const int64_t lo = P.ranges[p], hi = P.ranges[p + 1];
            std::vector<int32_t> ids(P.order.begin() + lo, P.order.begin() + hi);  // original ids of the part's nodes
            std::vector<int64_t> sz;
            rcb(xyz, dim, ids, 0, ids.size(), n2, sz);
            size_t o = 0;
            for (int a = 0; a < n2; ++a) {
                const int g = p * n2 + a;
                for (int64_t k = 0; k < sz[(size_t)a]; ++k, ++o) {
                    const int32_t orig = ids[o];
                    P.agg2[(size_t)P.inv[orig]] = g;
                    for (int d = 0; d < dim; ++d) P.cen2[(size_t)g * dim + d] += xyz[(size_t)orig * dim + d];
                }
                for (int d = 0; d < dim; ++d) P.cen2[(size_t)g * dim + d] /= (double)std::max<int64_t>(1, sz[(size_t)a]);
            }
        }
    }
    P.xyz.resize((size_t)n_nodes * dim);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_nodes; ++k)
        for (int d = 0; d < dim; ++d) P.xyz[(size_t)k * dim + d] = xyz[(size_t)P.order[k] * dim + d];
    // ---- 2. the mesh in the new numbering
    auto renumber = [&](const int32_t* conn, int64_t n, int npe, std::vector<int32_t>& out) -> bool {
        out.resize((size_t)n * npe);
        bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
        for (int64_t q = 0; q < n * npe; ++q) {
            const int32_t nd = conn[q];
            if (nd < 0 || nd >= n_nodes) {
                bad = true;
                out[q] = 0;
            } else
                out[q] = P.inv[nd];
        }
        return !bad;
    };
    if (!renumber(tets, n_tets, 4, P.tets) || !renumber(trusses, n_trusses, 2, P.trusses))
        return "element references a node id out of range";
    P.tet_has_mat = tet_mat != nullptr;
    if (tet_mat) P.tet_mat.assign(tet_mat, tet_mat + n_tets);
    P.truss_has_mat = truss_mat != nullptr;
    if (truss_mat) P.truss_mat.assign(truss_mat, truss_mat + n_trusses);
    if (n_trusses > 0 && area) P.area.assign(area, area + n_trusses);
    P.free_mask.assign((size_t)n_nodes * dim, 0);
    for (int64_t k = 0; k < n_free; ++k) {
        const int64_t g = free_dofs[k];
        if (g < 0 || g >= n_nodes * dim) return "free dof out of range";
        P.free_mask[(size_t)P.inv[g / dim] * dim + (size_t)(g % dim)] = 1;
    }
    P.n_free = n_free;
    // ---- 3. which ranks need which node: one pass over the elements
    P.need.assign((size_t)n_nodes, 0);
    std::vector<uint8_t> own((size_t)n_nodes);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_nodes; ++k) own[k] = (uint8_t)P.owner_of(k);
    auto mark = [&](const std::vector<int32_t>& conn, int64_t n, int npe) {
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < n; ++e) {
            unsigned ranks = 0;
            for (int b = 0; b < npe; ++b) ranks |= 1u << own[conn[(size_t)e * npe + b]];
            if ((ranks & (ranks - 1)) == 0) continue;  // all nodes on one rank
            for (int b = 0; b < npe; ++b) {
                const int32_t nd = conn[(size_t)e * npe + b];
                const uint16_t others = (uint16_t)(ranks & ~(1u << own[nd]));
                if ((P.need[nd] & others) != others) __atomic_fetch_or(&P.need[nd], others, __ATOMIC_RELAXED);
            }
        }
    };
    mark(P.tets, n_tets, 4);
    mark(P.trusses, n_trusses, 2);
    P.halo_cnt.assign((size_t)n_ranks * n_ranks, 0);
    for (int64_t k = 0; k < n_nodes; ++k) {
        unsigned m = P.need[k];
        while (m) {
            const int r = __builtin_ctz(m);
            m &= m - 1;
            P.halo_cnt[(size_t)r * n_ranks + own[k]]++;
        }
    }
    return std::string();
}

void build_local_part(const Partition& P, int rank, LocalPart& L) {
    L = LocalPart();
    const int N = P.n_ranks, dim = P.dim;
    const int64_t lo = P.ranges[rank], hi = P.ranges[rank + 1];
    L.rank = rank;
    L.n_ranks = N;
    L.dim = dim;
    L.n_owned = hi - lo;
    // ---- halo: nodes some element of mine touches, owned elsewhere; ascending new id = grouped by owner
    std::vector<int32_t> halo;
    const uint16_t mybit = (uint16_t)(1u << rank);
    for (int r = 0; r < N; ++r) {
        if (r == rank) continue;
        if (P.halo_cnt[(size_t)rank * N + r] == 0) continue;
        for (int64_t k = P.ranges[r]; k < P.ranges[r + 1]; ++k)
            if (P.need[k] & mybit) halo.push_back((int32_t)k);
    }
    L.n_local = L.n_owned + (int64_t)halo.size();
    L.l2g.resize((size_t)L.n_local);
    L.xyz.resize((size_t)L.n_local * dim);
    auto put = [&](int64_t li, int64_t gnew) {
        L.l2g[li] = P.order[gnew];
        for (int d = 0; d < dim; ++d) L.xyz[(size_t)li * dim + d] = P.xyz[(size_t)gnew * dim + d];
    };
#pragma omp parallel for schedule(static)
    for (int64_t k = lo; k < hi; ++k) put(k - lo, k);
    for (size_t h = 0; h < halo.size(); ++h) put(L.n_owned + (int64_t)h, halo[h]);
    auto to_local = [&](int32_t g) -> int32_t {
        if (g >= lo && g < hi) return (int32_t)(g - lo);
        return (int32_t)(L.n_owned + (std::lower_bound(halo.begin(), halo.end(), g) - halo.begin()));
    };
    // ---- neighbours, receive ranges (halo grouped by owner), send lists, remote offsets
    for (int r = 0; r < N; ++r)
        if (r != rank && (P.halo_cnt[(size_t)rank * N + r] > 0 || P.halo_cnt[(size_t)r * N + rank] > 0)) L.nbr_rank.push_back(r);
    const int nn = (int)L.nbr_rank.size();
    L.recv_ptr.assign((size_t)nn + 1, 0);
    L.send_ptr.assign((size_t)nn + 1, 0);
    L.remote_halo_off.assign((size_t)nn, 0);
    for (int k = 0; k < nn; ++k) {
        const int r = L.nbr_rank[k];
        L.recv_ptr[k + 1] = L.recv_ptr[k] + P.halo_cnt[(size_t)rank * N + r];
        const uint16_t rbit = (uint16_t)(1u << r);
        for (int64_t g = lo; g < hi; ++g)
            if (P.need[g] & rbit) L.send_nodes.push_back((int32_t)(g - lo));
        L.send_ptr[k + 1] = (int64_t)L.send_nodes.size();
        int64_t off = 0;  // r's halo is grouped by owner: my block starts after the blocks of the owners before me
        for (int o = 0; o < rank; ++o) off += P.halo_cnt[(size_t)r * N + o];
        L.remote_halo_off[k] = off;
    }
    // ---- elements touching an owned node, ascending global id (local order = global order: owned rows are summed in
    //      the same order as on one GPU)
    auto cut = [&](const std::vector<int32_t>& conn, int64_t n, int npe, std::vector<int32_t>& lconn, std::vector<int64_t>& gid) {
        std::vector<uint8_t> take((size_t)n, 0);
        int64_t cnt = 0;
#pragma omp parallel for schedule(static) reduction(+ : cnt)
        for (int64_t e = 0; e < n; ++e) {
            bool t = false;
            for (int b = 0; b < npe; ++b) {
                const int32_t g = conn[(size_t)e * npe + b];
                t = t || (g >= lo && g < hi);
            }
            take[e] = t ? 1 : 0;
            cnt += t ? 1 : 0;
        }
        gid.clear();
        gid.reserve((size_t)cnt);
        for (int64_t e = 0; e < n; ++e)
            if (take[e]) gid.push_back(e);
        lconn.resize((size_t)cnt * npe);
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < cnt; ++k)
            for (int b = 0; b < npe; ++b) lconn[(size_t)k * npe + b] = to_local(conn[(size_t)gid[k] * npe + b]);
    };
    cut(P.tets, P.n_tets, 4, L.tets, L.tet_global);
    cut(P.trusses, P.n_trusses, 2, L.trusses, L.truss_global);
    L.tet_mat.resize(L.tet_global.size(), 0);
    if (P.tet_has_mat)
        for (size_t k = 0; k < L.tet_global.size(); ++k) L.tet_mat[k] = P.tet_mat[(size_t)L.tet_global[k]];
    L.truss_mat.resize(L.truss_global.size(), 0);
    if (P.truss_has_mat)
        for (size_t k = 0; k < L.truss_global.size(); ++k) L.truss_mat[k] = P.truss_mat[(size_t)L.truss_global[k]];
    L.area.resize(L.truss_global.size(), 1.0);
    if (!P.area.empty())
        for (size_t k = 0; k < L.truss_global.size(); ++k) L.area[k] = P.area[(size_t)L.truss_global[k]];
    // ---- free dofs of the owned nodes, local numbering
    for (int64_t g = lo; g < hi; ++g)
        for (int d = 0; d < dim; ++d)
            if (P.free_mask[(size_t)g * dim + d]) L.free_dofs.push_back((g - lo) * dim + d);
    for (size_t h = 0; h < halo.size(); ++h)  // and of the halo nodes (after the owned ones): the solver updates U there too
        for (int d = 0; d < dim; ++d)
            if (P.free_mask[(size_t)halo[h] * dim + d]) L.free_dofs.push_back((L.n_owned + (int64_t)h) * dim + d);
    L.n_free_global = P.n_free;
    if ((size_t)rank < P.agg_ptr.size()) L.agg_ptr = P.agg_ptr[(size_t)rank];
    if (!P.agg2.empty()) {
        L.n_agg2_per_rank = P.n_agg2_per_rank;
        L.cen2 = P.cen2;
        L.agg2.resize((size_t)L.n_local);
        for (int64_t k = lo; k < hi; ++k) L.agg2[(size_t)(k - lo)] = P.agg2[(size_t)k];
        for (size_t h = 0; h < halo.size(); ++h) L.agg2[(size_t)L.n_owned + h] = P.agg2[(size_t)halo[h]];
    }
}

}  // namespace onsas
