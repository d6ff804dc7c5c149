// partition.hpp -- element-wise domain decomposition of a global mesh over the GPUs of one box.
//
// The reference has no distributed path (a single process loops over all elements, StaticAnalyses.jl:105-118); this
// is new.  Plain C++ (no CUDA): onsas_finalize_mesh of a multi-device context and the one-process-per-GPU binding
// (onsas_part_*) both call it, and tests/hostsim exercises it on the CPU.
//
//   * recursive coordinate bisection of the NODES into n_ranks parts; nodes are renumbered so that rank p owns the
//     contiguous range ranges[p] .. ranges[p+1], and inside a part the original order is kept (a structured mesh
//     stays x-fastest: the row-owner pair lists and the SpMV gathers stay local);
//   * rank p evaluates every element that touches one of its owned nodes (interface elements on both sides: no
//     assembly communication, the rows of K are complete on their owner);
//   * its halo = the other nodes of those elements, grouped by owner, ascending global id inside a group;
//   * what neighbour r needs from p = p's owned nodes that share an element with one of r's owned nodes.
// One pass over the elements marks, per node, the set of ranks (other than its owner) that need it (a 16-bit mask:
// at most 16 ranks); everything else -- halo lists, send lists, the offsets of p's values inside r's halo -- follows
// from that mask, so a process that only wants ITS part (one process per GPU) does no work for the others.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace onsas {

constexpr int PART_MAX_RANKS = 16;
// size of the coarse space of the two-level preconditioner (shared with the solver): at most COARSE_NC_MAX coarse dofs
// per device (the dense inverse, 18.9 MB, stays in L2 and fits the Gauss-Jordan kernel's shared memory), aggregates of
// about COARSE_TARGET_NODES nodes (7^3) when the mesh is large enough
constexpr int COARSE_NC_MAX = 1536;
constexpr int COARSE_TARGET_NODES = 343;
int coarse_aggregate_count(int64_t n_owned_nodes, int dim);

struct Partition {
    int n_ranks = 0, dim = 3;
    int64_t n_nodes = 0, n_tets = 0, n_trusses = 0;
    std::vector<int32_t> order;    // [n_nodes] new id -> original id
    std::vector<int32_t> inv;      // [n_nodes] original id -> new id
    std::vector<int64_t> ranges;   // [n_ranks+1] rank p owns new ids ranges[p] .. ranges[p+1]
    std::vector<uint16_t> need;    // [n_nodes] (new ids) bit r: rank r != owner has the node in its halo
    std::vector<int64_t> halo_cnt; // [n_ranks*n_ranks] halo_cnt[r*n_ranks + o] = nodes owned by o in the halo of r
    // the global mesh in the NEW numbering (kept: local parts are cut from it on demand)
    std::vector<double> xyz;       // [n_nodes*dim]
    std::vector<int32_t> tets, tet_mat, trusses, truss_mat;
    bool tet_has_mat = false, truss_has_mat = false;
    std::vector<double> area;
    std::vector<std::vector<int32_t>> agg_ptr;  // reorder = 2: per rank, the node ranges (relative to the rank's first node) of its aggregates
    // global coarse level of the two-level preconditioner (n_ranks > 1): every rank's nodes are cut into n_agg2_per_rank
    // "level-2" aggregates (recursive coordinate bisection, numbering untouched); their union over the ranks is ONE coarse
    // space of at most COARSE_NC_MAX dofs that couples the ranks
    std::vector<int32_t> agg2;     // [n_nodes] (new ids) global level-2 aggregate of every node; empty when n_ranks == 1
    std::vector<double> cen2;      // [n_agg2 * dim] centroids
    int n_agg2_per_rank = 0;
    std::vector<uint8_t> free_mask;  // [n_nodes*dim] (new numbering) 1 = free dof
    int64_t n_free = 0;
    int owner_of(int64_t new_id) const;
};

struct LocalPart {
    int rank = 0, n_ranks = 1, dim = 3;
    int64_t n_owned = 0, n_local = 0;
    std::vector<int32_t> l2g;       // [n_local] ORIGINAL global node id of each local node: owned first, then halo by owner
    std::vector<double> xyz;        // [n_local*dim]
    std::vector<int32_t> tets, tet_mat, trusses, truss_mat;   // local node ids
    std::vector<int64_t> tet_global, truss_global;            // global element id of each local element (ascending)
    std::vector<double> area;
    std::vector<int64_t> free_dofs; // local dofs that are free: those of the owned nodes first, then those of the halo nodes
    int64_t n_free_global = 0;
    std::vector<int32_t> nbr_rank;  // neighbours, ascending
    std::vector<int64_t> send_ptr, recv_ptr;   // [n_nbr+1]
    std::vector<int32_t> send_nodes;           // local (owned) node ids grouped by neighbour, ascending global id
    std::vector<int64_t> remote_halo_off;      // [n_nbr] where this rank's values start inside neighbour k's halo (in nodes)
    std::vector<int32_t> agg_ptr;              // reorder = 2: [n_agg+1] owned-node ranges of the preconditioner's aggregates, else empty
    std::vector<int32_t> agg2;                 // [n_local] global level-2 aggregate of every local node (owned + halo), or empty
    std::vector<double> cen2;                  // [n_agg2_total * dim] centroids of all level-2 aggregates
    int n_agg2_per_rank = 0;                   // this rank owns level-2 aggregates rank * n_agg2_per_rank .. + n_agg2_per_rank
};

// Returns an empty string on success.  conn arrays are element-major, 0-based original node ids; mat ids may be null.
std::string build_partition(int dim, int64_t n_nodes, const double* xyz, int64_t n_tets, const int32_t* tets,
                            const int32_t* tet_mat, int64_t n_trusses, const int32_t* trusses, const int32_t* truss_mat,
                            const double* area, int64_t n_free, const int64_t* free_dofs, int n_ranks, int reorder, Partition& out);
// reorder: 0 = inside a part the caller's node order is kept (meshes numbered for locality: structured grids, banded
// numberings); 1 = inside a part the nodes follow the Z-curve of their coordinates (any numbering becomes local);
// 2 = aggregate-major: the part is cut into the aggregates of the two-level preconditioner, numbered one after the other
// (Z-curve inside each), and LocalPart::agg_ptr hands their ranges to the solver.

void build_local_part(const Partition& P, int rank, LocalPart& out);

}  // namespace onsas
