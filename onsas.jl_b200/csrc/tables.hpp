// tables.hpp -- host-side preprocessing done once per mesh by onsas_finalize_mesh():
// node->element "pair" lists for the row-owner assembly, the block sparsity pattern of K in
// the sliced-ELL layout the SpMV reads, and the fixed-order contribution lists that make the
// assembly deterministic.  Replaces the reference's per-iteration COO buffers and per-entry
// sparse insertion (StructuralSolvers/Assemblers.jl:16-88, StaticAnalyses.jl:125-132).
//
// Layout ("BSELL-C", C = SLICE_ROWS block rows per slice):
//   slice sl covers block rows [sl*C, sl*C+C); width[sl] = max #blocks of its rows;
//   slice_ptr = exclusive prefix sum of width (units: block-columns);
//   col[(slice_ptr[sl] + s)*C + lane]                 -> node id of the s-th block of row sl*C+lane
//   val[((slice_ptr[sl] + s)*BS*BS + k)*C + lane]     -> entry k = BS*r + c of that block
//   padding blocks have col = the row itself and zero values.
// Within a row the blocks are sorted by node id, so block (row, col) is found by binary search.
//
// A "pair" is (row node i, element e, local node a with conn[e][a] == i); the thread that
// evaluates it produces the block-row a of K_e and f_a.  Pairs of a row are sorted by element
// id, so every entry of K and F_int is summed in ascending element order (the reference's
// order for a single-material structure) -- run-to-run and launch-configuration independent.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace onsas {

constexpr int SLICE_ROWS = 8;
constexpr int ROW_SKEW = 2;  // bank spreading of the shared-memory pair records: the records of row l of a slice are shifted by
// row_skew(l) doubles.  Rows are skewed in PAIRS: on the structured tet mesh (24 pairs per row) every other row boundary
// falls inside a half-warp of phase A, whose 16 stores stay conflict-free only if both rows have the same skew; a bank
// model of phase A stores + phase B loads on interior slices gives 1680 wavefronts per slice without skew, 1293 for
// l, 1161 for 2 l and 1005 for 2 (l / 2) (phase A at its ideal 468).
// The truss records (21 doubles, 14 pairs per row on the braced lattice) need none: the same model gives 294 wavefronts
// per slice without skew against 423 / 432 for l / 2 (l / 2).  family: 0 = tets, 1 = trusses.
constexpr inline int row_skew(int family, int l) { return family == 0 ? ROW_SKEW * (l >> 1) : 0; }
constexpr int TET_REC = 39;  // shared-memory record of one (row, tet) pair: 4 blocks * 9 + 3 force entries (odd stride)
constexpr inline int truss_rec(int dim) { return (2 * dim * dim + dim) | 1; }

// Everything a CTA needs to know about its slice, fetched with one 48-byte read.
struct alignas(16) SliceHdr {
    int64_t pair_base;                  // first pair of the slice in pair_nodes / pair_code
    int64_t slot_base;                  // slice_ptr[sl]: first block column of the slice
    int32_t n_pairs;
    int32_t width;                      // blocks per row of the slice (padded)
    uint16_t row_off[SLICE_ROWS + 1];   // pair range of each row, relative to pair_base
    uint16_t n_snodes;                  // distinct nodes the slice's elements touch (staged in shared memory by the CTA)
    uint32_t snode_base;                // first entry of the slice in snodes
};
static_assert(sizeof(SliceHdr) == 48, "SliceHdr must be 48 bytes");

struct FamilyTables {
    int npe = 0;                     // nodes per element (4 tets, 2 trusses)
    int rec = 0;                     // doubles per pair record in shared memory
    int64_t n_elem = 0;
    std::vector<int64_t> pair_ptr;   // [n_rows+1] pairs of row i
    std::vector<int32_t> pair_code;  // [n_pairs] e*npe + a, ascending e within a row
    std::vector<uint16_t> pair_lnodes; // [n_pairs*npe] the element's nodes as indices into the slice's node list; bit 15 of the
                                     // first one: this pair writes the element's stress / strain record (the pair of the
                                     // element's first OWNED node, so every local element is written on every rank)
    std::vector<int32_t> snodes;     // per slice: the distinct (local) node ids its elements touch, ascending
    int32_t max_snodes = 0;          // largest such list
    std::vector<SliceHdr> hdr;       // [n_slices]
    std::vector<uint32_t> cptr;      // [n_slots+1] contribution ranges per block slot (slot = slice_ptr*C + s*C + lane)
    std::vector<uint16_t> ccode;     // [n_pairs*npe] shared-memory offset of the contributing block: local_pair*rec + b*dim*dim
    int32_t max_pairs_per_slice = 0;
};

struct MeshTables {
    int dim = 3;
    int64_t n_nodes = 0;  // local nodes (owned + halo)
    int64_t n_rows = 0;   // owned nodes = block rows of K
    int64_t n_slices = 0;
    int32_t max_width = 0;           // widest slice (blocks per row)
    std::vector<int64_t> slice_ptr;  // [n_slices+1]
    std::vector<int32_t> col;        // [slice_ptr.back()*C]
    std::vector<int32_t> row_nblk;   // [n_rows] true number of blocks per row
    int64_t nnz_blocks = 0;          // sum of row_nblk
    FamilyTables fam[2];             // 0 = tets, 1 = trusses
    int64_t n_slots() const { return slice_ptr.empty() ? 0 : slice_ptr.back() * SLICE_ROWS; }
};

// conn arrays are element-major (npe per element), 0-based local node ids.
// Returns empty string on success, else an error message.
std::string build_mesh_tables(int dim, int64_t n_nodes, int64_t n_rows, int64_t n_tets, const int32_t* tets,
                              int64_t n_trusses, const int32_t* trusses, MeshTables& out);

// Slice ranges for a pipelined assembly with host buffers (onsas_assemble_host): `chunks` ranges, the first and the last
// of weight 1 against `mid_weight` for the inner ones; slice0[k] .. slice0[k+1] are the slices of range k and the elements
// evaluated by those slices touch the OWNED nodes [0, node_hi[k]) only (node_hi is non-decreasing, so U can travel as
// prefixes); halo nodes (multi-GPU: ids >= n_rows) are outside this accounting -- their block of U is sent first.
void host_range_plan(const MeshTables& t, int chunks, int mid_weight, std::vector<int64_t>& slice0, std::vector<int64_t>& node_hi);

// Scalar CSR (n_rows*dim rows, n_nodes*dim columns, sorted columns) index arrays of the same pattern.
void bsell_to_csr_pattern(const MeshTables& t, std::vector<int64_t>& rowptr, std::vector<int32_t>& colidx);
// Scatter BSELL values into the CSR value array produced for the pattern above.
void bsell_to_csr_values(const MeshTables& t, const double* val, double* csr_val);

}  // namespace onsas
